// EnCodec decoder kernels (24 kHz model, bandwidth 6 -> 8 codebooks, hop 320).
//
// Replaces encodec_forward_quantizer_decode (encodec.cpp/quantizer.h:78-111) and
// encodec_forward_decoder (encodec.cpp/decoder.h:43-113) with direct CUDA kernels:
//   RVQ gather-sum -> conv k7 -> 2 x LSTM(512) + skip -> 4 x [ELU, ConvT(k=2s, s), resblock] -> ELU -> conv k7.
// Activations are [C][T] with time contiguous (the reference's [T, C] ggml tensors).  Operands
// are rounded to f16 where the reference rounds them (im2col, ggml.c:14954; conv_transpose_1d,
// ggml.c:14659; LSTM mul_mat src1 conversion) and accumulated in f32; the summation order is NOT
// the reference's — the contract for the waveform is 1e-3 relative (BASELINE.json), not bit parity.
#include "codec_kernels.h"

namespace bark {

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }           // ggml.c:2533

// quantizer decode: x[d][t] = sum_q embed_q[codes[q][t]][d], q = 0..7 in order onto a zeroed tensor
struct Codebooks { const float * e[8]; };
__global__ void rvq_decode_kernel(Codebooks cb, const int32_t * __restrict__ codes, int T, int Hd, float * __restrict__ x) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y;
    if (t >= T) return;
    float acc = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) acc = __fadd_rn(acc, cb.e[q][(size_t) codes[q * T + t] * Hd + d]);
    x[(size_t) d * T + t] = acc;
}

// causal conv1d, stride 1 (ops.cpp:59-75): reflect pad k-1 on the left, y = b + sum_{c,j} w[o][c][j] * f16(x[c][t+j-(k-1)])
// optional ELU on the input (the decoder applies it right before most convs), optional residual add on the output.
template <int KW>
__global__ void conv1d_kernel(const float * __restrict__ x, int Cin, int T, const __half * __restrict__ w, const float * __restrict__ bias,
                              int Cout, int elu_in, const float * __restrict__ resid, float * __restrict__ y) {
    extern __shared__ float xs[];                        // [Cin][TILE + KW - 1], already ELU'd and f16-rounded
    const int TILE = blockDim.x;
    const int t0 = blockIdx.x * TILE;
    const int span = TILE + KW - 1;
    for (int i = threadIdx.x; i < Cin * span; i += blockDim.x) {
        const int c = i / span, j = i % span;
        int t = t0 + j - (KW - 1);
        if (t < 0) t = -t;                               // reflect (ggml.c:15589)
        float v = 0.f;
        if (t < T) { v = x[(size_t) c * T + t]; if (elu_in) v = elu1(v); v = round_f16(v); }
        xs[i] = v;
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= T) return;
    for (int o = blockIdx.y; o < Cout; o += gridDim.y) {
        const __half * wo = w + (size_t) o * Cin * KW;
        float acc = 0.f;
        for (int c = 0; c < Cin; c++) {
#pragma unroll
            for (int j = 0; j < KW; j++) acc = fmaf(__half2float(wo[c * KW + j]), xs[c * span + threadIdx.x + j], acc);
        }
        acc += bias[o];
        if (resid) acc += resid[(size_t) o * T + t];
        y[(size_t) o * T + t] = acc;
    }
}

void conv1d(const float * x, int Cin, int T, const ConvW & cv, bool elu_in, const float * resid, float * y, cudaStream_t s) {
    const int TILE = 128;
    const size_t smem = (size_t) Cin * (TILE + cv.k - 1) * sizeof(float);
    const dim3 grid((T + TILE - 1) / TILE, min(cv.cout, 64));
#define CONV_CASE(KW)                                                                                              \
    case KW:                                                                                                       \
        BARK_CUDA_CHECK(cudaFuncSetAttribute(conv1d_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
        BARK_LAUNCH(conv1d_kernel<KW>, grid, TILE, smem, s, x, Cin, T, cv.w, cv.b, cv.cout, elu_in ? 1 : 0, resid, y);   \
        break;
    switch (cv.k) { CONV_CASE(1) CONV_CASE(3) CONV_CASE(7) default: fprintf(stderr, "bark_b200: unsupported conv kernel size %d\n", cv.k); abort(); }
#undef CONV_CASE
}

// transposed conv (ops.cpp:77-98, ggml.c:14614-14700): k = 2*stride, right-trimmed by k - stride -> L = T*stride.
// y[o][p] = b[o] + sum_c ( w[c][o][p - t1*s] * f16(elu(x[c][t1])) + w[c][o][p - t0*s] * f16(elu(x[c][t0])) ), t1 = p/s - 1, t0 = p/s
__global__ void convtr1d_kernel(const float * __restrict__ x, int Cin, int T, const __half * __restrict__ w, const float * __restrict__ bias,
                                int Cout, int stride, float * __restrict__ y) {
    extern __shared__ float xs[];                        // [Cin][FR + 1] input frames t_lo-1 .. t_lo+FR-1
    const int FR = blockDim.x / stride;                  // frames per block
    const int tf0 = blockIdx.x * FR;
    for (int i = threadIdx.x; i < Cin * (FR + 1); i += blockDim.x) {
        const int c = i / (FR + 1), j = i % (FR + 1);
        const int t = tf0 + j - 1;
        xs[i] = (t >= 0 && t < T) ? round_f16(elu1(x[(size_t) c * T + t])) : 0.f;
    }
    __syncthreads();
    const int L = T * stride, K = 2 * stride;
    const int pl = threadIdx.x;                          // local output sample
    const int p = tf0 * stride + pl;
    if (pl >= FR * stride || p >= L) return;
    const int f = pl / stride, j0 = pl % stride;         // frame t0 = tf0+f uses tap j0, frame t0-1 uses tap j0+stride
    for (int o = blockIdx.y; o < Cout; o += gridDim.y) {
        float a1 = 0.f, a0 = 0.f;
        for (int c = 0; c < Cin; c++) {
            const __half * wc = w + ((size_t) c * Cout + o) * K;
            a1 = fmaf(__half2float(wc[j0 + stride]), xs[c * (FR + 1) + f], a1);        // earlier frame first (accumulation order of ggml.c:14688-14699)
            a0 = fmaf(__half2float(wc[j0]), xs[c * (FR + 1) + f + 1], a0);
        }
        y[(size_t) o * L + p] = bias[o] + (a1 + a0);
    }
}

void convtr1d(const float * x, int Cin, int T, const ConvW & cv, int stride, float * y, cudaStream_t s) {
    const int FR = max(1, 160 / stride);
    const int threads = FR * stride;
    const size_t smem = (size_t) Cin * (FR + 1) * sizeof(float);
    BARK_LAUNCH(convtr1d_kernel, dim3((T + FR - 1) / FR, min(cv.cout, 64)), threads, smem, s, x, Cin, T, cv.w, cv.b, cv.cout, stride, y);
}

// LSTM input projection for all time steps: gi[t][g] = b_ih[g] + W_ih[g][:] . f16(x[:][t])
__global__ void lstm_inproj_kernel(const float * __restrict__ x, int C, int T, const __half * __restrict__ wih, const float * __restrict__ bih,
                                   int G4, float * __restrict__ gi) {
    extern __shared__ float xs[];                        // [C] one time step
    const int t = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) xs[c] = round_f16(x[(size_t) c * T + t]);
    __syncthreads();
    for (int g = threadIdx.x; g < G4; g += blockDim.x) {
        const __half2 * wr = reinterpret_cast<const __half2 *>(wih + (size_t) g * C);
        float acc = 0.f;
        for (int c = 0; c < C / 2; c++) { const float2 wv = __half22float2(wr[c]); acc = fmaf(wv.x, xs[2 * c], acc); acc = fmaf(wv.y, xs[2 * c + 1], acc); }
        gi[(size_t) t * G4 + g] = acc + bih[g];
    }
}

// LSTM recurrence (lstm.h:52-73), one CTA walks the sequence; W_hh streams from L2 every step.
// out[j][t] (+= skip[j][t] when given).  Gate order i, f, g, o.
__global__ void __launch_bounds__(1024) lstm_recur_kernel(const float * __restrict__ gi, int T, int Hn, const __half * __restrict__ whh,
                                                          const float * __restrict__ bhh, const float * __restrict__ skip, float * __restrict__ out) {
    extern __shared__ float sm[];
    float * h16 = sm;                 // [Hn] h rounded to f16
    float * gates = sm + Hn;          // [4*Hn]
    const int G4 = 4 * Hn;
    float c_state = 0.f;              // thread j < Hn owns unit j
    for (int j = threadIdx.x; j < Hn; j += blockDim.x) h16[j] = 0.f;
    __syncthreads();
    for (int t = 0; t < T; t++) {
        for (int g = threadIdx.x; g < G4; g += blockDim.x) {
            const __half2 * wr = reinterpret_cast<const __half2 *>(whh + (size_t) g * Hn);
            float a0 = 0.f, a1 = 0.f;
            for (int c = 0; c < Hn / 2; c++) { const float2 wv = __half22float2(wr[c]); a0 = fmaf(wv.x, h16[2 * c], a0); a1 = fmaf(wv.y, h16[2 * c + 1], a1); }
            gates[g] = gi[(size_t) t * G4 + g] + ((a0 + a1) + bhh[g]);
        }
        __syncthreads();
        if ((int) threadIdx.x < Hn) {
            const int j = threadIdx.x;
            const float it = 1.f / (1.f + expf(-gates[j]));
            const float ft = 1.f / (1.f + expf(-gates[Hn + j]));
            const float gt = tanhf(gates[2 * Hn + j]);
            const float ot = 1.f / (1.f + expf(-gates[3 * Hn + j]));
            c_state = ft * c_state + it * gt;
            const float h = ot * tanhf(c_state);
            h16[j] = round_f16(h);
            out[(size_t) j * T + t] = skip ? h + skip[(size_t) j * T + t] : h;
        }
        __syncthreads();
    }
}

void lstm_layer(const float * x, int C, int T, const __half * wih, const __half * whh, const float * bih, const float * bhh,
                const float * skip, float * gi_scratch, float * out, cudaStream_t s) {
    const int Hn = C, G4 = 4 * Hn;
    BARK_LAUNCH(lstm_inproj_kernel, T, 512, (size_t) C * sizeof(float), s, x, C, T, wih, bih, G4, gi_scratch);
    BARK_LAUNCH(lstm_recur_kernel, 1, 1024, (size_t) 5 * Hn * sizeof(float), s, gi_scratch, T, Hn, whh, bhh, skip, out);
}

void rvq_decode(const CodecModel & cm, const int32_t * d_codes, int T, float * x, cudaStream_t s) {
    Codebooks cb; for (int q = 0; q < 8; q++) cb.e[q] = cm.embed[q];
    BARK_LAUNCH(rvq_decode_kernel, dim3((T + 127) / 128, cm.hidden_dim), 128, 0, s, cb, d_codes, T, cm.hidden_dim, x);
}

}  // namespace bark
