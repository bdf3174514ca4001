// EnCodec decoder kernels (24 kHz model, bandwidth 6 -> 8 codebooks, hop 320), bit-exact path.
//
// Replaces encodec_forward_quantizer_decode (encodec.cpp/quantizer.h:78-111) and
// encodec_forward_decoder (encodec.cpp/decoder.h:43-113):
//   RVQ gather-sum -> conv k7 -> 2 x LSTM(512) + skip -> 4 x [ELU, ConvT(k=2s, s), resblock] -> ELU -> conv k7.
// Activations are [C][T] with time contiguous (the reference's [T, C] ggml tensors).
//
// Every contraction in the reference's decoder is a ggml_vec_dot_f16 (ggml.c:2251): conv1d = im2col to f16
// (ggml.c:14892-14960) x f16 kernel, conv_transpose_1d = per-tap dots over Cin (ggml.c:14688-14699), LSTM = two
// f16 mat-vecs per step (lstm.h:55-59).  They are evaluated here in the same lane order as the GPT mat-muls
// (common.cuh), with the activation functions restated from glibc (common.cuh), so the waveform comes out
// bit-identical to the CPU reference, not merely within the 1e-3 contract.
#include "codec_kernels.h"

namespace bark {

// ------------------------------------------------------------------------------------------------
// quantizer decode: x[d][t] = sum_q embed_q[codes[q][t]][d], q = 0..7 in order onto a zeroed tensor
// ------------------------------------------------------------------------------------------------
struct Codebooks { const float * e[8]; };
__global__ void rvq_decode_kernel(Codebooks cb, const int32_t * __restrict__ codes, int T, int Hd, float * __restrict__ x) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y;
    if (t >= T) return;
    float acc = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) acc = __fadd_rn(acc, cb.e[q][(size_t) codes[q * T + t] * Hd + d]);
    x[(size_t) d * T + t] = acc;
}

void rvq_decode(const CodecModel & cm, const int32_t * d_codes, int T, float * x, cudaStream_t s) {
    Codebooks cb; for (int q = 0; q < 8; q++) cb.e[q] = cm.embed[q];
    BARK_LAUNCH(rvq_decode_kernel, dim3((T + 127) / 128, cm.hidden_dim), 128, 0, s, cb, d_codes, T, cm.hidden_dim, x);
}

// chain of one virtual lane out of an LI16 row: up to NG groups of 8 f16 values
template <int NG>
__device__ __forceinline__ void load_chain(const __half * __restrict__ row, int lane, int ngroups, float (&w)[NG * 8]) {
#pragma unroll
    for (int g = 0; g < NG; g++) {
        if (g < ngroups) {
            const uint4 u = __ldg(reinterpret_cast<const uint4 *>(row) + g * 32 + lane);
            const __half2 * h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
            for (int i = 0; i < 4; i++) { const float2 f = __half22float2(h[i]); w[g * 8 + 2 * i] = f.x; w[g * 8 + 2 * i + 1] = f.y; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// causal conv1d, stride 1 (ops.cpp:59-75): reflect-pad k-1 on the left (ggml.c:15589), im2col rounds the input to
// f16, y[o][t] = b[o] + vec_dot_f16(Cin*k, col[t], w[o]) with col index c*k + j.  Optional ELU on the input
// (decoder.h applies it right before most convs) and residual add on the output (decoder.h:101).
// One block: a tile of TT output positions x a chunk of output channels; the input tile sits in shared memory
// already ELU'd and f16-rounded; a warp keeps one filter's lane chains in registers and walks its positions.
// ------------------------------------------------------------------------------------------------
template <int KW, int NG>
__global__ void __launch_bounds__(256) conv1d_lane_kernel(const float * __restrict__ x, int Cin, int T, const __half * __restrict__ w_li, int Kp,
                                                          const float * __restrict__ bias, int Cout, int o_per_block, int elu_in,
                                                          const float * __restrict__ resid, float * __restrict__ y) {
    constexpr int TT = 32;
    constexpr int S = ((TT + KW - 1) | 1);               // odd row stride: conflict-free lane -> (c, j) gathers
    extern __shared__ float xs[];                        // [Cin][S]
    const int t0 = blockIdx.x * TT;
    for (int i = threadIdx.x; i < Cin * (TT + KW - 1); i += blockDim.x) {
        const int c = i / (TT + KW - 1), j = i % (TT + KW - 1);
        int t = t0 + j - (KW - 1);
        if (t < 0) t = -t;
        float v = 0.f;
        if (t < T) { v = x[(size_t) c * T + t]; if (elu_in) v = elu_exact(v); v = round_f16(v); }
        xs[c * S + j] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = Cin * KW, nsteps = K >> 5, ngroups = (nsteps + 7) >> 3;
    int aoff[NG * 8];                                    // smem offset of this lane's chain elements (position-independent part)
#pragma unroll
    for (int c = 0; c < NG * 8; c++) { const int kk = c * 32 + lane; aoff[c] = (kk / KW) * S + (kk % KW); }
    const int o_lo = blockIdx.y * o_per_block, o_hi = min(Cout, o_lo + o_per_block);
    for (int o = o_lo + warp; o < o_hi; o += 8) {
        float wch[NG * 8];
        load_chain<NG>(w_li + (size_t) o * Kp, lane, ngroups, wch);
        const float bo = bias[o];
        for (int tl = 0; tl < TT; tl++) {
            const int t = t0 + tl;
            if (t >= T) break;
            float acc = 0.0f;
#pragma unroll
            for (int c = 0; c < NG * 8; c++) if (c < nsteps) acc = __fmaf_rn(wch[c], xs[aoff[c] + tl], acc);
            float r = lane_tree_reduce(acc);
            if (lane == 0) {
                if ((nsteps << 5) < K) {                                         // K % 32 leftovers: float products summed in double (ggml.c:2281-2283)
                    double sd = (double) r;
                    const __half * wrow = w_li + (size_t) o * Kp;
                    for (int kk = nsteps << 5; kk < K; kk++)
                        sd = __dadd_rn(sd, (double) __fmul_rn(__half2float(wrow[li_offset(kk, 8)]), xs[(kk / KW) * S + (kk % KW) + tl]));
                    r = __double2float_rn(sd);
                }
                r = __fadd_rn(bo, r);                                            // ops.cpp:72 add(repeat(b), dst)
                if (resid) r = __fadd_rn(r, resid[(size_t) o * T + t]);
                y[(size_t) o * T + t] = r;
            }
        }
    }
}

void conv1d(const float * x, int Cin, int T, const ConvW & cv, bool elu_in, const float * resid, float * y, cudaStream_t s) {
    const int K = Cin * cv.k, nsteps = K / 32, ngroups = (nsteps + 7) / 8;
    if (ngroups > 4) { fprintf(stderr, "bark_b200: unsupported conv shape Cin=%d k=%d\n", Cin, cv.k); throw std::runtime_error("unsupported configuration (see the message above)"); }
    const int TT = 32;
    const int S = (TT + cv.k - 1) | 1;
    const size_t smem = (size_t) Cin * S * sizeof(float);
    const int tiles = (T + TT - 1) / TT;
    // enough blocks to fill the machine: split the output channels when there are few time tiles
    int o_per_block = cv.cout;
    while (o_per_block > 8 && tiles * ((cv.cout + o_per_block - 1) / o_per_block) < 4 * 148) o_per_block = (o_per_block + 1) / 2;
    const dim3 grid(tiles, (cv.cout + o_per_block - 1) / o_per_block);
    g_next_flops = 2.0 * (double) T * cv.cout * K;
#define CONV_CASE(KW, NG)                                                                                                   \
    { BARK_CUDA_CHECK(cudaFuncSetAttribute(conv1d_lane_kernel<KW, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); \
      BARK_LAUNCH((conv1d_lane_kernel<KW, NG>), grid, 256, smem, s, x, Cin, T, cv.w, cv.Kp, cv.b, cv.cout, o_per_block, elu_in ? 1 : 0, resid, y); }
    if (cv.k == 1)      { if (ngroups <= 1) CONV_CASE(1, 1) else CONV_CASE(1, 2) }
    else if (cv.k == 3) { if (ngroups <= 1) CONV_CASE(3, 1) else if (ngroups <= 2) CONV_CASE(3, 2) else CONV_CASE(3, 3) }
    else if (cv.k == 7) { if (ngroups <= 1) CONV_CASE(7, 1) else CONV_CASE(7, 4) }
    else { fprintf(stderr, "bark_b200: unsupported conv kernel size %d\n", cv.k); throw std::runtime_error("unsupported configuration (see the message above)"); }
#undef CONV_CASE
}

// ------------------------------------------------------------------------------------------------
// transposed conv (ops.cpp:77-98, ggml.c:14614-14700): k = 2*stride, output right-trimmed by k - stride -> L = T*stride.
// For output sample p = t*stride + j (0 <= j < stride) the reference accumulates, in this order,
//     v1 = vec_dot_f16(Cin, f16(elu(x[:, t-1])), w[:, o, j + stride])     (skipped for t = 0)
//     v0 = vec_dot_f16(Cin, f16(elu(x[:, t])),   w[:, o, j])
// into a zeroed buffer, then adds the bias.  Weights arrive re-laid-out as rows [o][tap][Cin] in LI16.
// ------------------------------------------------------------------------------------------------
template <int NG>
__global__ void __launch_bounds__(256) convtr1d_lane_kernel(const float * __restrict__ x, int Cin, int T, const __half * __restrict__ w_li, int Kp,
                                                            const float * __restrict__ bias, int Cout, int stride, float * __restrict__ y) {
    constexpr int TF = 16;                               // input frames per block
    constexpr int S = TF + 1 + ((TF + 1) % 2 == 0);      // odd stride
    extern __shared__ float xs[];                        // [Cin][S]: frames t0-1 .. t0+TF-1, ELU'd, f16-rounded
    const int t0 = blockIdx.x * TF;
    for (int i = threadIdx.x; i < Cin * (TF + 1); i += blockDim.x) {
        const int c = i / (TF + 1), j = i % (TF + 1);
        const int t = t0 + j - 1;
        xs[c * S + j] = (t >= 0 && t < T) ? round_f16(elu_exact(x[(size_t) c * T + t])) : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nsteps = Cin >> 5, ngroups = (nsteps + 7) >> 3;
    const int K2 = 2 * stride, L = T * stride;
    const int items = Cout * stride;                     // work items: (o, j) pairs, spread over blockIdx.y and the warps
    for (int it = blockIdx.y * 8 + warp; it < items; it += gridDim.y * 8) {
        const int o = it / stride, j = it % stride;
        float w0[NG * 8], w1[NG * 8];
        load_chain<NG>(w_li + ((size_t) o * K2 + j) * Kp, lane, ngroups, w0);
        load_chain<NG>(w_li + ((size_t) o * K2 + j + stride) * Kp, lane, ngroups, w1);
        const float bo = bias[o];
        for (int f = 0; f < TF; f++) {
            const int t = t0 + f;
            if (t >= T) break;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int c = 0; c < NG * 8; c++) if (c < nsteps) {
                const float * col = xs + (c * 32 + lane) * S + f;
                a1 = __fmaf_rn(col[0], w1[c], a1);       // frame t-1, tap j+stride
                a0 = __fmaf_rn(col[1], w0[c], a0);       // frame t,   tap j
            }
            const float r1 = lane_tree_reduce(a1), r0 = lane_tree_reduce(a0);
            if (lane == 0) {
                float acc = 0.0f;
                if (t > 0) acc = __fadd_rn(acc, r1);
                acc = __fadd_rn(acc, r0);
                y[(size_t) o * L + (size_t) t * stride + j] = __fadd_rn(bo, acc);
            }
        }
    }
}

void convtr1d(const float * x, int Cin, int T, const ConvW & cv, int stride, float * y, cudaStream_t s) {
    const int nsteps = Cin / 32, ngroups = (nsteps + 7) / 8;
    if (Cin % 32 != 0 || ngroups > 2 || cv.k != 2 * stride) { fprintf(stderr, "bark_b200: unsupported transposed conv Cin=%d k=%d s=%d\n", Cin, cv.k, stride); throw std::runtime_error("unsupported configuration (see the message above)"); }
    const int TF = 16, S = TF + 1 + ((TF + 1) % 2 == 0);
    const size_t smem = (size_t) Cin * S * sizeof(float);
    const int tiles = (T + TF - 1) / TF;
    int gy = (cv.cout * stride + 7) / 8;
    while (gy > 1 && tiles * gy > 8 * 148) gy = (gy + 1) / 2;
    g_next_flops = 2.0 * 2.0 * (double) T * stride * cv.cout * Cin;
    if (ngroups <= 1) BARK_LAUNCH(convtr1d_lane_kernel<1>, dim3(tiles, gy), 256, smem, s, x, Cin, T, cv.w, cv.Kp, cv.b, cv.cout, stride, y);
    else              BARK_LAUNCH(convtr1d_lane_kernel<2>, dim3(tiles, gy), 256, smem, s, x, Cin, T, cv.w, cv.Kp, cv.b, cv.cout, stride, y);
}

// ------------------------------------------------------------------------------------------------
// LSTM (lstm.h:22-78).  Input projections for all steps at once, then the recurrence.
//   gates[t] = (W_ih f16(x_t) + b_ih) + (W_hh f16(h_{t-1}) + b_hh);  i,f,o = sigmoid, g = tanh (order i,f,g,o)
//   c = f*c + i*g;  h = o * tanh(c)
// ------------------------------------------------------------------------------------------------
// gi[t][g] = vec_dot_f16(C, w_ih[g], f16(x[:, t])) + b_ih[g];   one warp per (t, gate row) pair, weights chain reused over 8 steps
__global__ void __launch_bounds__(256) lstm_inproj_lane_kernel(const float * __restrict__ x, int C, int T, const __half * __restrict__ wih_li, int Kp,
                                                               const float * __restrict__ bih, int G4, float * __restrict__ gi) {
    constexpr int TT = 8;
    extern __shared__ float xs[];                        // [TT][C] f16-rounded
    const int t0 = blockIdx.x * TT;
    for (int i = threadIdx.x; i < TT * C; i += blockDim.x) {
        const int tl = i / C, c = i % C;
        xs[i] = (t0 + tl < T) ? round_f16(x[(size_t) c * T + t0 + tl]) : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nsteps = C >> 5;                           // 16 for C = 512
    for (int g = blockIdx.y * 8 + warp; g < G4; g += gridDim.y * 8) {
        float wch[16];
        load_chain<2>(wih_li + (size_t) g * Kp, lane, (nsteps + 7) >> 3, wch);
        const float bg = bih[g];
        for (int tl = 0; tl < TT && t0 + tl < T; tl++) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 16; c++) if (c < nsteps) acc = __fmaf_rn(wch[c], xs[tl * C + c * 32 + lane], acc);
            const float r = lane_tree_reduce(acc);
            if (lane == 0) gi[(size_t)(t0 + tl) * G4 + g] = __fadd_rn(r, bg);
        }
    }
}

// Recurrence: persistent cooperative kernel.  CTA b owns UPB hidden units (4*UPB gate rows of W_hh, held in registers by its
// warps for the whole sequence); every step each CTA reads h_{t-1} (written by all CTAs), computes its gates, updates its
// units and publishes h_t, then all CTAs meet at a grid barrier (monotonic counter in global memory).
__device__ __forceinline__ void grid_barrier(unsigned * counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*((volatile unsigned *) counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

template <int UPB>
__global__ void __launch_bounds__(UPB * 4 * 32) lstm_recur_kernel(const float * __restrict__ gi, int T, int Hn, const __half * __restrict__ whh_li, int Kp,
                                                                  const float * __restrict__ bhh, const float * __restrict__ skip,
                                                                  float * __restrict__ hbuf /*[2][Hn]*/, unsigned * __restrict__ counter, float * __restrict__ out) {
    extern __shared__ float hs[];                        // [Hn] f16-rounded h_{t-1}; then [4*UPB] gate pre-activations
    float * gates = hs + Hn;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;       // warp = local gate row: unit u = warp % UPB, gate q = warp / UPB
    const int u = warp % UPB, q = warp / UPB;
    const int unit = blockIdx.x * UPB + u;
    const int row = q * Hn + unit;
    const int G4 = 4 * Hn, nsteps = Hn >> 5;
    float wch[16];
    load_chain<2>(whh_li + (size_t) row * Kp, lane, (nsteps + 7) >> 3, wch);
    const float bg = bhh[row];
    float c_state = 0.f;                                 // meaningful in thread (warp = u, lane 0)... kept by threads 0..UPB-1 instead
    for (int t = 0; t < T; t++) {
        const float * hprev = hbuf + (size_t)((t + 1) & 1) * Hn;
        for (int j = threadIdx.x; j < Hn; j += blockDim.x) hs[j] = (t == 0) ? 0.f : round_f16(__ldcg(hprev + j));
        __syncthreads();
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 16; c++) if (c < nsteps) acc = __fmaf_rn(wch[c], hs[c * 32 + lane], acc);
        const float r = lane_tree_reduce(acc);
        if (lane == 0) gates[warp] = __fadd_rn(__ldg(gi + (size_t) t * G4 + row), __fadd_rn(r, bg));     // (ih + b_ih) + (hh + b_hh)
        __syncthreads();
        if ((int) threadIdx.x < UPB) {
            const int uu = threadIdx.x, un = blockIdx.x * UPB + uu;
            const float it = sigmoid_exact(gates[0 * UPB + uu]);
            const float ft = sigmoid_exact(gates[1 * UPB + uu]);
            const float gt = glibc_tanhf_dev(gates[2 * UPB + uu]);
            const float ot = sigmoid_exact(gates[3 * UPB + uu]);
            c_state = __fadd_rn(__fmul_rn(ft, c_state), __fmul_rn(it, gt));
            const float h = __fmul_rn(ot, glibc_tanhf_dev(c_state));
            __stcg(hbuf + (size_t)(t & 1) * Hn + un, h);
            out[(size_t) un * T + t] = skip ? __fadd_rn(skip[(size_t) un * T + t], h) : h;               // decoder.h:72 inpL + out
        }
        grid_barrier(counter, (unsigned)(t + 1) * gridDim.x);
    }
}

void lstm_layer(const float * x, int C, int T, const __half * wih_li, const __half * whh_li, int Kp, const float * bih, const float * bhh,
                const float * skip, float * gi_scratch, float * hbuf, unsigned * counter, float * out, cudaStream_t s) {
    const int Hn = C, G4 = 4 * Hn;
    if (Hn % 32 != 0 || Hn > 512 || Hn % 4 != 0) { fprintf(stderr, "bark_b200: unsupported LSTM width %d\n", Hn); throw std::runtime_error("unsupported configuration (see the message above)"); }
    g_next_flops = 2.0 * (double) T * G4 * C;
    BARK_LAUNCH(lstm_inproj_lane_kernel, dim3((T + 7) / 8, 32), 256, (size_t) 8 * C * sizeof(float), s, x, C, T, wih_li, Kp, bih, G4, gi_scratch);
    BARK_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned), s));
    constexpr int UPB = 4;
    const int blocks = Hn / UPB;                         // 128 CTAs for H = 512: co-resident on 148 SMs (cooperative launch checks it)
    const size_t smem = (size_t)(Hn + 4 * UPB) * sizeof(float);
    void * args[] = {(void *) &gi_scratch, (void *) &T, (void *) &Hn, (void *) &whh_li, (void *) &Kp, (void *) &bhh, (void *) &skip, (void *) &hbuf, (void *) &counter, (void *) &out};
    if (g_prof_on) prof_begin("lstm_recur_kernel", s, 0.0, 2.0 * (double) T * G4 * Hn);
    BARK_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *) lstm_recur_kernel<UPB>, dim3(blocks), dim3(UPB * 4 * 32), args, smem, s));
    if (g_prof_on) prof_end(s);
    ++g_kernel_launches;
}

// [Cin][Cout][k] (torch ConvTranspose1d layout as stored, ggml [k, Cout, Cin]) -> rows [o][tap][Cin]
__global__ void convtr_rows_kernel(const __half * __restrict__ src, __half * __restrict__ dst, int Cin, int Cout, int k) {
    const size_t total = (size_t) Cin * Cout * k;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const int c = (int)(i % Cin), j = (int)((i / Cin) % k), o = (int)(i / ((size_t) Cin * k));
        dst[i] = src[((size_t) c * Cout + o) * k + j];
    }
}
void convtr_rows(const __half * src, __half * dst, int Cin, int Cout, int k, cudaStream_t s) {
    BARK_LAUNCH(convtr_rows_kernel, 592, 256, 0, s, src, dst, Cin, Cout, k);
}

}  // namespace bark
