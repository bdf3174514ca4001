// Register-tiled lane-order kernels for the multi-row passes: prefill of the causal models (257 / <=887 rows) and the
// fine model's 1024-row passes (bark.cpp:1416-1584) — the dense contractions of the hot path.
//
// Bit-exactness fixes the shape of the tiling: every output element is 32 lane partials (common.cuh "Lane order"),
// each a serial FMA chain over k = v, v+32, ..., so a warp's 32 lanes all work on the SAME outputs — lane v owns
// virtual lane v of an 8 x 8 output tile (64 accumulators per lane, 2048 per warp).  Operand reuse therefore comes
// from registers (each converted operand feeds 8 FMAs) and shared memory (a 32 x 16 block tile), not from giving
// different lanes different outputs.  The 32 partials of all 64 outputs are then combined with a transposed
// butterfly (62 shuffles instead of 64 x 5) that performs exactly the additions of GGML_F32x8_REDUCE.
//
// Both operands are in the lane-interleaved layout, so a k-step of the block tile is 48 rows x 512 contiguous
// bytes: one TMA bulk copy per row into a 4-stage shared-memory ring (mbarrier complete_tx).
#include "epilogue.cuh"
#include "gpt_kernels.h"

namespace bark {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void * src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Combine the 32 lane partials of 64 outputs.  In: r[i] = this lane's partial of output i.  Out: r[0], r[1] = the complete
// sums of outputs base, base+1 with base = (b4<<5)|(b3<<4)|(b2<<3)|(b0<<2)|(b1<<1), b = bits of the lane id.
// Stage order xor 16, 8, 4, 1, 2 = the tree of GGML_F32x8_REDUCE (ggml.c:1405-1422); each add is commutative.
template <int N, int MASK>
__device__ __forceinline__ void butterfly_stage(float (&r)[64], bool upper) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        const float keep = upper ? r[i + N / 2] : r[i];
        const float send = upper ? r[i] : r[i + N / 2];
        r[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, MASK));
    }
}
__device__ __forceinline__ int butterfly_reduce64(float (&r)[64], int lane) {
    butterfly_stage<64, 16>(r, (lane & 16) != 0);
    butterfly_stage<32, 8>(r, (lane & 8) != 0);
    butterfly_stage<16, 4>(r, (lane & 4) != 0);
    butterfly_stage<8, 1>(r, (lane & 1) != 0);
    butterfly_stage<4, 2>(r, (lane & 2) != 0);
    return ((lane & 16) << 1) | ((lane & 8) << 1) | ((lane & 4) << 1) | ((lane & 1) << 2) | (lane & 2);
}

constexpr int kBM = 32, kBO = 16, kStages = 4, kRowBytes = 512;
constexpr int kStageBytes = (kBM + kBO) * kRowBytes;     // 24 KB

template <typename T> struct Cvt;
template <> struct Cvt<__half> {
    static constexpr int G = 8;
    __device__ __forceinline__ static float elem(const uint4 & u, int e) {        // e is a compile-time constant after unrolling
        const uint32_t w = e < 2 ? u.x : e < 4 ? u.y : e < 6 ? u.z : u.w;
        const __half2 h = *reinterpret_cast<const __half2 *>(&w);
        return (e & 1) ? __high2float(h) : __low2float(h);
    }
};
template <> struct Cvt<float> {
    static constexpr int G = 4;
    __device__ __forceinline__ static float elem(const uint4 & u, int e) { return __uint_as_float(e == 0 ? u.x : e == 1 ? u.y : e == 2 ? u.z : u.w); }
};

}  // namespace

// C[m][o] = lane-order dot(act[m], W[o]).  Persistent CTAs walk 32 x 16 block tiles (o fastest, so CTAs running at the
// same time share activation rows in L2); 8 warps as 4 (m) x 2 (o), 8 x 8 outputs per warp.  The k-steps of ALL of a CTA's
// tiles form one stream through the 4-stage ring: warp 0 lane-issues the bulk copies of step s + 4 as soon as the 8 warps
// have released the slot (per-slot `empty` mbarrier), so the loads of the next tile fly while this tile's tail and
// butterfly epilogue run — with K = 768 a tile is only 3 k-steps, and a non-persistent version spent a third of its
// time filling the pipe (ncu: 37% FMA-pipe active, top stall = barrier).
template <typename T>
__global__ void __launch_bounds__(256, 2) lane_gemm_tiled_kernel(const T * __restrict__ W, int K, int Kp, int O, const T * __restrict__ act, int M, MatmulEpilogue ep) {
    constexpr int G = Cvt<T>::G;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t full = smem_u32(smem + kStages * kStageBytes), empty = full + kStages * 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wo = warp & 1;
    const int nsteps = K >> 5;
    const int ngroups = (nsteps + G - 1) / G;                 // k-steps per tile; the last may be partial
    const int tiles_o = (O + kBO - 1) / kBO, tiles_m = (M + kBM - 1) / kBM, n_tiles = tiles_o * tiles_m;
    const int my_tiles = ((int) blockIdx.x < n_tiles) ? (n_tiles - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0;
    const int total_steps = my_tiles * ngroups;

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(full + s * 8, 1); mbar_init(empty + s * 8, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int step) {                               // executed by warp 0: global k-step `step` of this CTA's stream
        const int ti = step / ngroups, g = step - ti * ngroups;
        const int tile = blockIdx.x + ti * gridDim.x;
        const int m0 = (tile / tiles_o) * kBM, o0 = (tile % tiles_o) * kBO;
        const int slot = step % kStages;
        const uint32_t bar = full + slot * 8;
        if (lane == 0) mbar_expect_tx(bar, kStageBytes);
        __syncwarp();
        const uint32_t dst = smem_u32(smem + (size_t) slot * kStageBytes);
        {   // activation rows (clamped: rows past M repeat the last row and are masked at the store)
            const int m = min(m0 + lane, M - 1);
            tma_bulk_g2s(dst + lane * kRowBytes, (const unsigned char *) (act + (size_t) m * Kp) + (size_t) g * kRowBytes, kRowBytes, bar);
        }
        if (lane < kBO) {
            const int o = min(o0 + lane, O - 1);
            tma_bulk_g2s(dst + (kBM + lane) * kRowBytes, (const unsigned char *) (W + (size_t) o * Kp) + (size_t) g * kRowBytes, kRowBytes, bar);
        }
    };
    if (warp == 0) for (int s0 = 0; s0 < kStages && s0 < total_steps; s0++) issue(s0);

    int step = 0;
    for (int ti = 0; ti < my_tiles; ti++) {
        const int tile = blockIdx.x + ti * gridDim.x;
        const int m0 = (tile / tiles_o) * kBM, o0 = (tile % tiles_o) * kBO;
        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; i++) acc[i] = 0.0f;
        for (int g = 0; g < ngroups; g++, step++) {
            const int slot = step % kStages;
            const uint32_t use = (uint32_t)(step / kStages);
            mbar_wait(full + slot * 8, use & 1);
            const unsigned char * st = smem + (size_t) slot * kStageBytes;
            const int steps = min(G, nsteps - g * G);          // chain steps present in this group
            uint4 pa[8];
#pragma unroll
            for (int mi = 0; mi < 8; mi++) pa[mi] = *reinterpret_cast<const uint4 *>(st + (wm * 8 + mi) * kRowBytes + lane * 16);
#pragma unroll
            for (int oh = 0; oh < 2; oh++) {
                uint4 pw[4];
#pragma unroll
                for (int oi = 0; oi < 4; oi++) pw[oi] = *reinterpret_cast<const uint4 *>(st + (kBM + wo * 8 + oh * 4 + oi) * kRowBytes + lane * 16);
#pragma unroll
                for (int e = 0; e < G; e++) {
                    if (e < steps) {                           // uniform across the block
                        float wf[4];
#pragma unroll
                        for (int oi = 0; oi < 4; oi++) wf[oi] = Cvt<T>::elem(pw[oi], e);
#pragma unroll
                        for (int mi = 0; mi < 8; mi++) {
                            const float af = Cvt<T>::elem(pa[mi], e);
#pragma unroll
                            for (int oi = 0; oi < 4; oi++) acc[mi * 8 + oh * 4 + oi] = __fmaf_rn(wf[oi], af, acc[mi * 8 + oh * 4 + oi]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + slot * 8);      // this warp is done with the slot
            if (warp == 0 && step + kStages < total_steps) {   // refill it for k-step `step + kStages` once all 8 warps released it
                mbar_wait(empty + slot * 8, use & 1);
                issue(step + kStages);
            }
        }
        const int base = butterfly_reduce64(acc, lane);        // outputs base, base+1 of the warp tile (index = mi*8 + oi)
        const int m = m0 + wm * 8 + (base >> 3), o = o0 + wo * 8 + (base & 7);
        if (m < M) {
            if (o < O) matmul_epilogue(ep, m, o, acc[0]);
            if (o + 1 < O) matmul_epilogue(ep, m, o + 1, acc[1]);
        }
    }
}

void lane_gemm_tiled(const DMat & W, const void * act, int rows, const MatmulEpilogue & ep, cudaStream_t s) {
    const size_t smem = (size_t) kStages * kStageBytes + 2 * kStages * 8 + 64;
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0; BARK_CUDA_CHECK(cudaGetDevice(&dev));
        BARK_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        BARK_CUDA_CHECK(cudaFuncSetAttribute(lane_gemm_tiled_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        BARK_CUDA_CHECK(cudaFuncSetAttribute(lane_gemm_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    }
    const int n_tiles = ((W.n_out + kBO - 1) / kBO) * ((rows + kBM - 1) / kBM);
    const int grid = min(n_tiles, 2 * n_sm);                   // persistent: two CTAs per SM (registers and shared memory allow exactly that)
    if (W.type == W_F16) BARK_LAUNCH((lane_gemm_tiled_kernel<__half>), grid, 256, smem, s, (const __half *) W.p, W.K, W.Kp, W.n_out, (const __half *) act, rows, ep);
    else                 BARK_LAUNCH((lane_gemm_tiled_kernel<float>), grid, 256, smem, s, (const float *) W.p, W.K, W.Kp, W.n_out, (const float *) act, rows, ep);
}

// ------------------------------------------------------------------------------------------------
// attention, multi-row (bark.cpp:1302-1339 / 1495-1530): 8 x 8 tiles per warp, same lane mapping
// ------------------------------------------------------------------------------------------------
// scores[h][q][k] = vec_dot_f32(D, K[k][h], Q[q][h]) * scale, -inf where k > n_past + q (causal)
template <int DSTEPS>
__global__ void __launch_bounds__(256) attn_scores_tiled_kernel(const float * __restrict__ Q, const float * __restrict__ Kc, int N, int n_kv, int n_past, int E,
                                                                float scale, int causal, float * __restrict__ S) {
    constexpr int D = DSTEPS * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.z, q0 = blockIdx.y * 8, k0 = (blockIdx.x * 8 + warp) * 8;
    if (k0 >= n_kv) return;
    float qf[8][DSTEPS], kf[8][DSTEPS];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int q = min(q0 + i, N - 1), k = min(k0 + i, n_kv - 1);
#pragma unroll
        for (int c = 0; c < DSTEPS; c++) {
            qf[i][c] = __ldg(Q + (size_t) q * E + h * D + c * 32 + lane);
            kf[i][c] = __ldg(Kc + (size_t) k * E + h * D + c * 32 + lane);
        }
    }
    float acc[64];
#pragma unroll
    for (int qi = 0; qi < 8; qi++)
#pragma unroll
        for (int ki = 0; ki < 8; ki++) {
            float a = 0.0f;
#pragma unroll
            for (int c = 0; c < DSTEPS; c++) a = __fmaf_rn(kf[ki][c], qf[qi][c], a);
            acc[qi * 8 + ki] = a;
        }
    const int base = butterfly_reduce64(acc, lane);
    const int q = q0 + (base >> 3), k = k0 + (base & 7);
    if (q < N) {
        float * row = S + ((size_t) blockIdx.z * N + q) * n_kv;
#pragma unroll
        for (int j = 0; j < 2; j++) if (k + j < n_kv) {
            float r = __fmul_rn(acc[j], scale);                                       // ggml_scale_inplace
            if (causal && k + j > n_past + q) r = __int_as_float(0xff800000);         // ggml_diag_mask_inf
            row[k + j] = r;
        }
    }
}

// KQV[q][h*D+d] = vec_dot_f32(n_kv, V^T[d], P[q]) -> activation operand for c_proj.  Warp = 8 queries x 8 head columns;
// lane v walks k = v, v+32, ...
__global__ void __launch_bounds__(256) attn_pv_tiled_kernel(const float * __restrict__ S, const float * __restrict__ Vc, int N, int n_kv, int E, int D,
                                                            void * __restrict__ act, int wt, int Kp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.y, q0 = blockIdx.x * 8, d0 = warp * 8;
    if (d0 >= D) return;
    const float * prow[8];
#pragma unroll
    for (int qi = 0; qi < 8; qi++) prow[qi] = S + ((size_t) h * N + min(q0 + qi, N - 1)) * n_kv;
    const float * vbase = Vc + h * D + d0;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; i++) acc[i] = 0.0f;
    const int np = n_kv & ~31;
    for (int k = lane; k < np; k += 32) {
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(vbase + (size_t) k * E));
        const float4 v1 = __ldg(reinterpret_cast<const float4 *>(vbase + (size_t) k * E) + 1);
        const float vf[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int qi = 0; qi < 8; qi++) {
            const float p = __ldg(prow[qi] + k);
#pragma unroll
            for (int di = 0; di < 8; di++) acc[qi * 8 + di] = __fmaf_rn(vf[di], p, acc[qi * 8 + di]);
        }
    }
    const int base = butterfly_reduce64(acc, lane);
    const int q = q0 + (base >> 3), d = d0 + (base & 7);
    if (q >= N) return;
    const float * p = S + ((size_t) h * N + q) * n_kv;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        float sum = acc[j];
        const float * v = Vc + h * D + d + j;
        int i = np, r = n_kv - np;                                                    // leftovers as compiled in the pinned build (orc_vec_dot_f32)
        while (r >= 8) { for (int l = 0; l < 8; l++) sum = __fadd_rn(sum, __fmul_rn(__ldg(v + (size_t)(i + l) * E), __ldg(p + i + l))); i += 8; r -= 8; }
        if (r >= 4)    { for (int l = 0; l < 4; l++) sum = __fadd_rn(sum, __fmul_rn(__ldg(v + (size_t)(i + l) * E), __ldg(p + i + l))); i += 4; r -= 4; }
        for (; r > 0; r--, i++) sum = __fmaf_rn(__ldg(v + (size_t) i * E), __ldg(p + i), sum);
        store_act(act, wt, Kp, q, h * D + d + j, sum);
    }
}

void attention_tiled_scores(const float * Q, const float * Kc, int N, int n_kv, int n_past, int E, int H, float scale, bool causal, float * scores, cudaStream_t s) {
    const int D = E / H;
    const dim3 grid((n_kv + 63) / 64, (N + 7) / 8, H);
    if (D == 64)       BARK_LAUNCH(attn_scores_tiled_kernel<2>, grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
    else if (D == 32)  BARK_LAUNCH(attn_scores_tiled_kernel<1>, grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
    else if (D == 96)  BARK_LAUNCH(attn_scores_tiled_kernel<3>, grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
    else               BARK_LAUNCH(attn_scores_tiled_kernel<4>, grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
}

void attention_tiled_pv(const float * scores, const float * Vc, int N, int n_kv, int E, int H, void * act, WType wt, int Kp, cudaStream_t s) {
    const int D = E / H;
    BARK_LAUNCH(attn_pv_tiled_kernel, dim3((N + 7) / 8, H), 32 * ((D + 7) / 8), 0, s, scores, Vc, N, n_kv, E, D, act, (int) wt, Kp);
}

}  // namespace bark
