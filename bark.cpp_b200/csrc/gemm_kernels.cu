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
// Both operands are in the group-major layout (common.cuh), so a pipeline stage of the block tile is a handful of
// contiguous spans: 2-4 TMA bulk copies into a 4-stage shared-memory ring (mbarrier complete_tx).
//
// F2 variants (the default since they were validated bit-exact on a B200, round 2; BARK_B200_FFMA2=0 selects the scalar-FMA
// kernels for A-B runs): the 64 independent FMAs of a chain step
// are issued as 32 packed FFMA2 (sm_100 `fma.rn.f32x2`, __ffma2_rn: "numeric behavior per component is the same as
// __fmaf_rn"), with the scalar operand broadcast by the instruction itself.  Same arithmetic, same order, half the issue
// slots — the tiled kernels are issue-bound (43 % of issue slots busy, 59-76 % of the instructions are FMAs).
#include "epilogue.cuh"
#include "gpt_kernels.h"

#include <string.h>

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

constexpr int kBM = 32, kBO = 16, kStages = 4;
constexpr int kStageBytes = (kBM + kBO) * 512;           // 24 KB: 256 columns of f16 (two groups) or 128 columns of f32 (one group)

// lane v's four elements of one row of one 128-column group (group-major layout, common.cuh)
template <typename T> struct QuadOp;
template <> struct QuadOp<__half> {
    typedef uint2 V;
    static constexpr int kGroupsPerStage = 2;
    __device__ __forceinline__ static float elem(const uint2 & u, int c) {       // c is a compile-time constant after unrolling
        const uint32_t w = c < 2 ? u.x : u.y;
        const __half2 h = *reinterpret_cast<const __half2 *>(&w);
        return (c & 1) ? __high2float(h) : __low2float(h);
    }
};
template <> struct QuadOp<float> {
    typedef uint4 V;
    static constexpr int kGroupsPerStage = 1;
    __device__ __forceinline__ static float elem(const uint4 & u, int c) { return __uint_as_float(c == 0 ? u.x : c == 1 ? u.y : c == 2 ? u.z : u.w); }
};

}  // namespace

// C[m][o] = lane-order dot(act[m], W[o]).  Persistent CTAs walk 32 x 16 block tiles (o fastest, so CTAs running at the
// same time share activation rows in L2); 8 warps as 4 (m) x 2 (o), 8 x 8 outputs per warp.
//
// Both operands are group-major: for one 128-column group the 32 activation rows of a tile are one contiguous span, the
// 16 weight rows another, so a pipeline stage is 2 (f32) or 4 (f16) bulk copies issued by ONE thread.  (A first version
// with row-major operands needed 48 row copies per stage; the copy instruction takes uniform operands, the compiler
// serialised the 48 lanes, and the issuing warp - also a consumer - fell ~1.5k cycles behind per stage: ncu showed 30% of
// all stall samples on the `full`-barrier wait of the other seven warps.)
//
// The k-steps of ALL of a CTA's tiles form one stream through the 4-stage ring.  Nobody waits to refill a slot: every warp
// bumps the slot's counter when it is done reading, and the warp that arrives last issues the copies for stage s + 4.
template <typename T, bool F2 = false>
__global__ void __launch_bounds__(256, 2) lane_gemm_tiled_kernel(const T * __restrict__ Wg, int K, int w_gs, int O, const T * __restrict__ act, int act_gs, int M, MatmulEpilogue ep) {
    typedef typename QuadOp<T>::V QV;
    constexpr int GPS = QuadOp<T>::kGroupsPerStage;
    constexpr int kRowB = kGmGroup * sizeof(T);              // bytes of one row of one group: 256 (f16) / 512 (f32)
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t full = smem_u32(smem + kStages * kStageBytes);
    int * const cnt = reinterpret_cast<int *>(smem + kStages * kStageBytes + kStages * 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wo = warp & 1;
    const int nsteps = K >> 5;                                // chain steps per lane
    const int ngroups = (nsteps + 3) >> 2;                    // 128-column groups; the last may be partial
    const int nstages = (ngroups + GPS - 1) / GPS;            // pipeline stages per tile
    const int tiles_o = (O + kBO - 1) / kBO, tiles_m = (M + kBM - 1) / kBM, n_tiles = tiles_o * tiles_m;
    const int my_tiles = ((int) blockIdx.x < n_tiles) ? (n_tiles - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0;
    const int total_steps = my_tiles * nstages;

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(full + s * 8, 1); cnt[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int step) {                              // ONE thread: stage `step` of this CTA's stream
        const int ti = step / nstages, sg = step - ti * nstages;
        const int tile = blockIdx.x + ti * gridDim.x;
        const int m0 = (tile / tiles_o) * kBM, o0 = (tile % tiles_o) * kBO;
        const int slot = step % kStages;
        const uint32_t bar = full + slot * 8;
        const int g0 = sg * GPS, ng = min(GPS, ngroups - g0);
        mbar_expect_tx(bar, (uint32_t)(ng * (kBM + kBO) * kRowB));
        const uint32_t dst = smem_u32(smem + (size_t) slot * kStageBytes);
        for (int j = 0; j < ng; j++) {
            tma_bulk_g2s(dst + j * (kBM + kBO) * kRowB, act + (size_t)(g0 + j) * act_gs + (size_t) m0 * kGmGroup, kBM * kRowB, bar);
            tma_bulk_g2s(dst + j * (kBM + kBO) * kRowB + kBM * kRowB, Wg + (size_t)(g0 + j) * w_gs + (size_t) o0 * kGmGroup, kBO * kRowB, bar);
        }
    };
    if (tid == 0) for (int s0 = 0; s0 < kStages && s0 < total_steps; s0++) issue(s0);

    int step = 0;
    for (int ti = 0; ti < my_tiles; ti++) {
        const int tile = blockIdx.x + ti * gridDim.x;
        const int m0 = (tile / tiles_o) * kBM, o0 = (tile % tiles_o) * kBO;
        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; i++) acc[i] = 0.0f;
        for (int sg = 0; sg < nstages; sg++, step++) {
            const int slot = step % kStages;
            mbar_wait(full + slot * 8, (uint32_t)(step / kStages) & 1);
            const unsigned char * st = smem + (size_t) slot * kStageBytes;
#pragma unroll
            for (int j = 0; j < GPS; j++) {
                const int g = sg * GPS + j;
                if (g < ngroups) {                            // uniform across the block
                    const int steps = min(4, nsteps - g * 4); // chain steps present in this group
                    const unsigned char * ga = st + j * (kBM + kBO) * kRowB + lane * sizeof(QV), * gw = ga + kBM * kRowB;
                    QV pa[8], pw[8];
#pragma unroll
                    for (int mi = 0; mi < 8; mi++) pa[mi] = *reinterpret_cast<const QV *>(ga + (wm * 8 + mi) * kRowB);
#pragma unroll
                    for (int oi = 0; oi < 8; oi++) pw[oi] = *reinterpret_cast<const QV *>(gw + (wo * 8 + oi) * kRowB);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        if (c < steps) {
                            float af[8], wf[8];
#pragma unroll
                            for (int mi = 0; mi < 8; mi++) af[mi] = QuadOp<T>::elem(pa[mi], c);
#pragma unroll
                            for (int oi = 0; oi < 8; oi++) wf[oi] = QuadOp<T>::elem(pw[oi], c);
                            if constexpr (F2) {
#pragma unroll
                                for (int mi = 0; mi < 8; mi++)
#pragma unroll
                                    for (int op = 0; op < 8; op += 2) {
                                        const float2 r = __ffma2_rn(make_float2(wf[op], wf[op + 1]), make_float2(af[mi], af[mi]), make_float2(acc[mi * 8 + op], acc[mi * 8 + op + 1]));
                                        acc[mi * 8 + op] = r.x; acc[mi * 8 + op + 1] = r.y;
                                    }
                            } else {
#pragma unroll
                            for (int mi = 0; mi < 8; mi++)
#pragma unroll
                                for (int oi = 0; oi < 8; oi++) acc[mi * 8 + oi] = __fmaf_rn(wf[oi], af[mi], acc[mi * 8 + oi]);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {                                  // release the slot; the last of the 8 warps refills it
                __threadfence_block();
                if (atomicAdd_block(&cnt[slot], 1) == 7) {
                    cnt[slot] = 0;
                    __threadfence_block();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (step + kStages < total_steps) issue(step + kStages);
                }
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

// packed-FMA variants of the three tiled kernels (see the header comment) unless BARK_B200_FFMA2=0
static bool use_ffma2() { const char * e = getenv("BARK_B200_FFMA2"); return !(e && e[0] == '0' && e[1] == 0); }   // read per launch so one process can A-B the two

void lane_gemm_tiled(const DMat & W, const void * act, int act_gs, int rows, const MatmulEpilogue & ep, cudaStream_t s, bool f32_containers) {
    const size_t smem = (size_t) kStages * kStageBytes + kStages * 8 + kStages * 4 + 64;
    int dev = 0, n_sm = 0;
    BARK_CUDA_CHECK(cudaGetDevice(&dev));
    BARK_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    static std::atomic<unsigned long long> configured{0};
    if (first_use_on_this_device(configured)) {
        BARK_CUDA_CHECK(cudaFuncSetAttribute(lane_gemm_tiled_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        BARK_CUDA_CHECK(cudaFuncSetAttribute(lane_gemm_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        BARK_CUDA_CHECK(cudaFuncSetAttribute((lane_gemm_tiled_kernel<__half, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        BARK_CUDA_CHECK(cudaFuncSetAttribute((lane_gemm_tiled_kernel<float, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    }
    if (f32_containers && W.type == W_F16) {                 // f16 values in f32 containers on both sides: the float kernel, bit-identical results
        if (!W.p_gm32) { fprintf(stderr, "bark_b200: matrix has no f32-expanded copy\n"); throw std::runtime_error("unsupported configuration (see the message above)"); }
        const int n_tiles32 = ((W.n_out + kBO - 1) / kBO) * ((rows + kBM - 1) / kBM);
        const int grid32 = min(n_tiles32, 2 * n_sm);
        if (use_ffma2()) BARK_LAUNCH((lane_gemm_tiled_kernel<float, true>), grid32, 256, smem, s, (const float *) W.p_gm32, W.K, W.o_pad * kGmGroup, W.n_out, (const float *) act, act_gs, rows, ep);
        else             BARK_LAUNCH((lane_gemm_tiled_kernel<float>), grid32, 256, smem, s, (const float *) W.p_gm32, W.K, W.o_pad * kGmGroup, W.n_out, (const float *) act, act_gs, rows, ep);
        return;
    }
    if (!W.p_gm) { fprintf(stderr, "bark_b200: matrix has no group-major copy for the tiled mat-mul\n"); throw std::runtime_error("unsupported configuration (see the message above)"); }
    const int n_tiles = ((W.n_out + kBO - 1) / kBO) * ((rows + kBM - 1) / kBM);
    const int grid = min(n_tiles, 2 * n_sm);                   // persistent: two CTAs per SM (registers and shared memory allow exactly that)
    const int w_gs = W.o_pad * kGmGroup;
    if (use_ffma2()) {
        if (W.type == W_F16) BARK_LAUNCH((lane_gemm_tiled_kernel<__half, true>), grid, 256, smem, s, (const __half *) W.p_gm, W.K, w_gs, W.n_out, (const __half *) act, act_gs, rows, ep);
        else                 BARK_LAUNCH((lane_gemm_tiled_kernel<float, true>), grid, 256, smem, s, (const float *) W.p_gm, W.K, w_gs, W.n_out, (const float *) act, act_gs, rows, ep);
        return;
    }
    if (W.type == W_F16) BARK_LAUNCH((lane_gemm_tiled_kernel<__half>), grid, 256, smem, s, (const __half *) W.p_gm, W.K, w_gs, W.n_out, (const __half *) act, act_gs, rows, ep);
    else                 BARK_LAUNCH((lane_gemm_tiled_kernel<float>), grid, 256, smem, s, (const float *) W.p_gm, W.K, w_gs, W.n_out, (const float *) act, act_gs, rows, ep);
}

// ------------------------------------------------------------------------------------------------
// attention, multi-row (bark.cpp:1302-1339 / 1495-1530): 8 x 8 tiles per warp, same lane mapping
// ------------------------------------------------------------------------------------------------
// scores[h][q][k] = vec_dot_f32(D, K[k][h], Q[q][h]) * scale, -inf where k > n_past + q (causal)
template <int DSTEPS, bool F2 = false>
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
    if constexpr (F2) {
#pragma unroll
        for (int qi = 0; qi < 8; qi++)
#pragma unroll
            for (int kp = 0; kp < 8; kp += 2) {
                float2 a = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int c = 0; c < DSTEPS; c++) a = __ffma2_rn(make_float2(kf[kp][c], kf[kp + 1][c]), make_float2(qf[qi][c], qf[qi][c]), a);
                acc[qi * 8 + kp] = a.x; acc[qi * 8 + kp + 1] = a.y;
            }
    } else {
#pragma unroll
    for (int qi = 0; qi < 8; qi++)
#pragma unroll
        for (int ki = 0; ki < 8; ki++) {
            float a = 0.0f;
#pragma unroll
            for (int c = 0; c < DSTEPS; c++) a = __fmaf_rn(kf[ki][c], qf[qi][c], a);
            acc[qi * 8 + ki] = a;
        }
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
template <bool F2 = false>
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
            if constexpr (F2) {
#pragma unroll
                for (int dp = 0; dp < 8; dp += 2) {
                    const float2 r = __ffma2_rn(make_float2(vf[dp], vf[dp + 1]), make_float2(p, p), make_float2(acc[qi * 8 + dp], acc[qi * 8 + dp + 1]));
                    acc[qi * 8 + dp] = r.x; acc[qi * 8 + dp + 1] = r.y;
                }
            } else {
#pragma unroll
            for (int di = 0; di < 8; di++) acc[qi * 8 + di] = __fmaf_rn(vf[di], p, acc[qi * 8 + di]);
            }
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
    if (use_ffma2()) {
        if (D == 64)       BARK_LAUNCH((attn_scores_tiled_kernel<2, true>), grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
        else if (D == 32)  BARK_LAUNCH((attn_scores_tiled_kernel<1, true>), grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
        else if (D == 96)  BARK_LAUNCH((attn_scores_tiled_kernel<3, true>), grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
        else               BARK_LAUNCH((attn_scores_tiled_kernel<4, true>), grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
        return;
    }
    if (D == 64)       BARK_LAUNCH(attn_scores_tiled_kernel<2>, grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
    else if (D == 32)  BARK_LAUNCH(attn_scores_tiled_kernel<1>, grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
    else if (D == 96)  BARK_LAUNCH(attn_scores_tiled_kernel<3>, grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
    else               BARK_LAUNCH(attn_scores_tiled_kernel<4>, grid, 256, 0, s, Q, Kc, N, n_kv, n_past, E, scale, causal ? 1 : 0, scores);
}

void attention_tiled_pv(const float * scores, const float * Vc, int N, int n_kv, int E, int H, void * act, WType wt, int Kp, cudaStream_t s) {
    const int D = E / H;
    if (use_ffma2()) { BARK_LAUNCH((attn_pv_tiled_kernel<true>), dim3((N + 7) / 8, H), 32 * ((D + 7) / 8), 0, s, scores, Vc, N, n_kv, E, D, act, (int) wt, Kp); return; }
    BARK_LAUNCH((attn_pv_tiled_kernel<false>), dim3((N + 7) / 8, H), 32 * ((D + 7) / 8), 0, s, scores, Vc, N, n_kv, E, D, act, (int) wt, Kp);
}

}  // namespace bark
