// Device-resident model description: what the loader produces and the kernels consume.
// Mirrors the reference's gpt_hparams / gpt_layer / gpt_model (bark.cpp:49-121) and encodec_model
// (encodec.cpp/encodec.cpp:47-99), but every pointer is HBM and every matrix is in the
// lane-interleaved layout of common.cuh.
#pragma once
#include "common.cuh"

#include <string>
#include <vector>

namespace bark {

struct DMat {                 // 2-D weight [n_out][K] in LI layout (f32 / f16) or q4_0 blocks
    void * p = nullptr;
    int n_out = 0, K = 0, Kp = 0;   // Kp = padded row length in elements
    void * scales = nullptr;        // quantised: f16 block scales [n_out][K/32]; p then holds the 16-byte nibble words (32 B for q8_0) [n_out][K/32]
    void * mins = nullptr, * qh = nullptr;   // experimental types: f16 block minima (q4_1, q5_1), fifth bits (q5_0, q5_1)
    void * p_gm = nullptr; int o_pad = 0;   // second copy in the group-major layout (common.cuh) for the tiled GEMM; rows padded to o_pad
    void * p_gm32 = nullptr;                // f16 matrices, BARK_B200_GEMM_F32C=1: the group-major copy expanded to f32 (gemm_kernels.cu)
    void * p_rm = nullptr;                  // fast mode only: the file's row-major [n_out][K] f16 matrix = K-major tcgen05 operand (fast_kernels.cu)
    WType type = W_F16;
};

struct GPTLayer {
    float * ln_1_g = nullptr, * ln_1_b = nullptr, * ln_2_g = nullptr, * ln_2_b = nullptr;
    DMat c_attn, c_proj, fc, proj;
};

struct GPTModel {
    // header order of the file (bark.cpp:700-709)
    int32_t n_layer = 0, n_head = 0, n_embd = 0, block_size = 0, bias = 0, n_in_vocab = 0, n_out_vocab = 0,
            n_lm_heads = 0, n_wtes = 0, ftype = 0;
    WType wtype = W_F16;
    void * wte[8] = {nullptr};        // token tables, ORIGINAL row-major layout (gather only)
    float * wpe = nullptr;            // [block_size][E] f32
    float * ln_f_g = nullptr, * ln_f_b = nullptr;
    DMat lm_head[8];
    std::vector<GPTLayer> layers;
    float * mem_k = nullptr, * mem_v = nullptr;   // [L][block_size][E] f32 (bark.cpp:980-981); null for the fine model
    // persistent decode step (decode_kernels.cu): phase table + cross-CTA exchange buffers, built once at load
    bool decode_ok = false;           // the model fits the persistent kernel's fixed capacities (build_decode_tables); otherwise per-op stepping
    void * d_phases = nullptr, * d_layer_vecs = nullptr;
    unsigned long long * gx = nullptr, * gq = nullptr, * gk = nullptr, * gv = nullptr, * gatt = nullptr, * gff = nullptr, * gscores = nullptr;
    float * glogits = nullptr;
    unsigned * d_adapt = nullptr;     // per-CTA adaptive head starts of the exchanges (decode_kernels.cu)
    // per-model statistics, same meaning as gpt_model::t_* (bark.cpp:114-118)
    int64_t t_sample_us = 0, t_predict_us = 0, t_main_us = 0, n_sample = 0;
};

struct ConvW { __half * w = nullptr; float * b = nullptr; int k = 0, cin = 0, cout = 0, Kp = 0; };   // w: LI16 rows (conv: [Cout] x Cin*k; transposed conv: [Cout*k] x Cin)

struct CodecModel {
    int hidden_dim = 128, n_filters = 32, kernel_size = 7, res_kernel = 3, n_bins = 1024;
    ConvW init, final_conv;
    __half * lstm_ih_w[2] = {nullptr, nullptr}, * lstm_hh_w[2] = {nullptr, nullptr};   // [4H] x H, LI16 rows
    int lstm_Kp = 0;
    float  * lstm_ih_b[2] = {nullptr, nullptr}, * lstm_hh_b[2] = {nullptr, nullptr};
    struct Block { ConvW us, c1, c2, sc; } blk[4];   // us: transposed conv
    float * embed[8] = {nullptr};           // codebooks 0..7, [n_bins][hidden] f32
};

// Scratch activations for one GPT evaluation of up to `max_rows` positions.
struct Workspace {
    int max_rows = 0, E = 0;
    float * x = nullptr;        // residual stream [rows][E]
    void  * act = nullptr;      // LI-layout activation operand for the next matmul [rows][max(Kp)]
    void  * act2 = nullptr;     // second operand buffer (GELU output feeding mlp/c_proj)
    float * q = nullptr;        // [rows][E]
    float * kbuf = nullptr, * vbuf = nullptr;   // fine model K/V [rows][E]
    float * scores = nullptr;   // [H][rows][n_kv]
    float * logits = nullptr;   // [rows][n_out]
    int32_t * tok = nullptr;    // device copy of the ids fed this step
};

}  // namespace bark
