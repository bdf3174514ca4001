/* Shim for the two things callers of bark.h take from ggml.h (reference: encodec.cpp/ggml/include/ggml.h):
 *   - enum ggml_ftype (ggml.h:388-417), used by bark_model_quantize's signature and examples/quantize;
 *   - ggml_time_init / ggml_time_us / ggml_time_ms (ggml.h:696-698), called by examples/main/main.cpp:26-27,81
 *     and examples/server/server.cpp:97-98.
 *   - ggml_init / ggml_free (ggml.h:640-660): examples/quantize/main.cpp:67-72 creates and frees an empty context "to initialize
 *     the f16 tables"; here that is a no-op returning a non-null token.
 * Nothing else of ggml exists in this library. */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
enum ggml_ftype {
    GGML_FTYPE_UNKNOWN = -1,
    GGML_FTYPE_ALL_F32 = 0, GGML_FTYPE_MOSTLY_F16 = 1, GGML_FTYPE_MOSTLY_Q4_0 = 2, GGML_FTYPE_MOSTLY_Q4_1 = 3,
    GGML_FTYPE_MOSTLY_Q4_1_SOME_F16 = 4, GGML_FTYPE_MOSTLY_Q8_0 = 7, GGML_FTYPE_MOSTLY_Q5_0 = 8, GGML_FTYPE_MOSTLY_Q5_1 = 9,
    GGML_FTYPE_MOSTLY_Q2_K = 10, GGML_FTYPE_MOSTLY_Q3_K = 11, GGML_FTYPE_MOSTLY_Q4_K = 12, GGML_FTYPE_MOSTLY_Q5_K = 13,
    GGML_FTYPE_MOSTLY_Q6_K = 14, GGML_FTYPE_MOSTLY_IQ2_XXS = 15, GGML_FTYPE_MOSTLY_IQ2_XS = 16, GGML_FTYPE_MOSTLY_IQ3_XXS = 17,
    GGML_FTYPE_MOSTLY_IQ1_S = 18, GGML_FTYPE_MOSTLY_IQ4_NL = 19, GGML_FTYPE_MOSTLY_IQ3_S = 20, GGML_FTYPE_MOSTLY_IQ2_S = 21,
    GGML_FTYPE_MOSTLY_IQ4_XS = 22, GGML_FTYPE_MOSTLY_IQ1_M = 23, GGML_FTYPE_MOSTLY_BF16 = 24, GGML_FTYPE_MOSTLY_Q4_0_4_4 = 25,
    GGML_FTYPE_MOSTLY_Q4_0_4_8 = 26, GGML_FTYPE_MOSTLY_Q4_0_8_8 = 27,
};
#include <stddef.h>
#include <stdbool.h>
struct ggml_context;
struct ggml_init_params { size_t mem_size; void * mem_buffer; bool no_alloc; };      /* ggml.h:640-645 */
__attribute__((visibility("default"))) struct ggml_context * ggml_init(struct ggml_init_params params);
__attribute__((visibility("default"))) void    ggml_free(struct ggml_context * ctx);
__attribute__((visibility("default"))) void    ggml_time_init(void);
__attribute__((visibility("default"))) int64_t ggml_time_ms(void);
__attribute__((visibility("default"))) int64_t ggml_time_us(void);
#ifdef __cplusplus
}
#endif
