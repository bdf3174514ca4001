// Device-side token sampling: gpt_sample (bark.cpp:184-270) for a batch of logit rows, bit-faithful to the host path.
//
// The reference samples on the CPU: l /= temp; float softmax with the DOUBLE exp() and a SEQUENTIAL float sum; then
// libstdc++'s std::discrete_distribution (double normalisation, sequential double partial sums, last one forced to 1,
// lower_bound of one generate_canonical<double,53> draw; bits/random.tcc:2657-2730).  Doing that on the host costs a
// logits read-back per step (4.3 MB per fine pass) plus 50-190 us of scalar work per sample — ~100 ms of a 2.76 s clip.
//
// Here one CTA owns one row:
//   * divisions, max, exps: parallel over the row;
//   * the float sum: strictly sequential on one thread (the order IS the result);
//   * exp: CUDA's double exp (<= 1 ulp) rounded to float.  glibc's exp is also < 1 ulp, so the two can only round to
//     different floats when the double lies within a few ulp of a float rounding boundary: that is detected
//     (both ends of a +-2^-50 relative bracket must round to the same float) and the row is FLAGGED;
//   * the multinomial draw: the uniform u comes from the host's std::mt19937 stream (same two 32-bit draws per sample);
//     the partial sums are formed in parallel in double and compared with u * total; if any partial sum lies within the
//     rounding-error bound of the threshold the row is FLAGGED.
// Flagged rows (probability ~1e-7 per row) are re-sampled on the host with the reference's exact sequence, using the same
// u, so the token stream is identical to the host path in all cases.
#include "gpt_kernels.h"
#include "sampling.cuh"

namespace bark {

// 256 threads per row when many rows are sampled at once (fine passes: 1024 rows), 1024 threads for the single row of a decode
// step (the exp / division / scan passes of a 10 048-wide semantic row are 4x shorter; the sequential sum is unchanged)

// One CTA per row.  tok_add is added to the sampled index (coarse stage: offset of the codebook window in
// the vocabulary); feed, when set, receives the token for the NEXT decode step to read (no host round trip).
template <int kSampleThreads>
__global__ void __launch_bounds__(kSampleThreads) sample_rows_kernel(const float * __restrict__ logits, int ld, int n, int rows, float temp, const double * __restrict__ u,
                                                                     int32_t * __restrict__ out_tok, int tok_add, int32_t * __restrict__ feed,
                                                                     float * __restrict__ eos_p, int32_t * __restrict__ flags, int force_flag) {
    extern __shared__ float sh[];                        // [n] working row
    const int row = blockIdx.x;
    sample_row_body<kSampleThreads>(sh, logits + (size_t) row * ld, n, temp, temp != 0.0f ? u[row] : 0.0, out_tok + row, tok_add, feed ? feed + row : nullptr,
                                    eos_p ? eos_p + row : nullptr, flags + row, force_flag);
}

void sample_rows(const float * logits, int ld, int n, int rows, float temp, const double * d_u, int32_t * d_out_tok, int tok_add, int32_t * d_feed,
                 float * d_eos_p, int32_t * d_flags, int force_flag, cudaStream_t s) {
    const size_t smem = ((size_t) n * sizeof(float) + 15) & ~(size_t) 15;
    static std::atomic<unsigned long long> configured{0};
    if (first_use_on_this_device(configured)) {
        BARK_CUDA_CHECK(cudaFuncSetAttribute(sample_rows_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        BARK_CUDA_CHECK(cudaFuncSetAttribute(sample_rows_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    }
    g_next_bytes = (double) rows * n * 4.0;
    if (rows == 1) BARK_LAUNCH(sample_rows_kernel<1024>, rows, 1024, smem, s, logits, ld, n, rows, temp, d_u, d_out_tok, tok_add, d_feed, d_eos_p, d_flags, force_flag);
    else           BARK_LAUNCH(sample_rows_kernel<256>, rows, 256, smem, s, logits, ld, n, rows, temp, d_u, d_out_tok, tok_add, d_feed, d_eos_p, d_flags, force_flag);
}

}  // namespace bark
