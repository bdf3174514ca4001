// Host-callable launchers of the EnCodec decoder kernels (codec_kernels.cu).
#pragma once
#include "model.h"

namespace bark {

void rvq_decode(const CodecModel & cm, const int32_t * d_codes /*[8][T]*/, int T, float * x /*[hidden][T]*/, cudaStream_t s);
void conv1d(const float * x, int Cin, int T, const ConvW & cv, bool elu_in, const float * resid, float * y, cudaStream_t s);
void convtr1d(const float * x, int Cin, int T, const ConvW & cv, int stride, float * y /*[Cout][T*stride]*/, cudaStream_t s);
void lstm_layer(const float * x, int C, int T, const __half * wih_li, const __half * whh_li, int Kp, const float * bih, const float * bhh,
                const float * skip, float * gi_scratch /*[T][4C]*/, float * hbuf /*[2][C]*/, unsigned * counter, float * out, cudaStream_t s);
// load-time re-layout of a transposed-conv weight: [Cin][Cout][k] -> rows [Cout][k][Cin]
void convtr_rows(const __half * src, __half * dst, int Cin, int Cout, int k, cudaStream_t s);

}  // namespace bark
