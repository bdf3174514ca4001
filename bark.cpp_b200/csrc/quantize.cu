// bark_model_quantize (bark.h:229-232): f32 / f16 ggml_weights.bin -> q4_0 ggml_weights.bin, host only.
//
// Same container walk as the reference (bark_model_quantize bark.cpp:2300-2377, bark_model_weights_quantize
// bark.cpp:2234-2298, ggml_quantize_weights bark.cpp:272-470): magic and vocabulary copied, each GPT section re-emitted with
// ftype = GGML_QNT_VERSION * 1000 + ftype, the 2-D tensors named wte / lm_head / c_attn / c_proj / c_fc / mlp c_proj
// quantised row by row with quantize_row_{q4_0,q4_1,q5_0,q5_1,q8_0}_ref (ggml-quants.c:668-871), everything else and the whole codec
// section copied byte for byte.  Output files are byte-identical to the reference tool's for all five types its README lists
// (tests/test_quantize.py).  The GPU path itself reads f32, f16 and q4_0 files; k-quants and i-quants are rejected with a message.
#include "../../include/bark.h"

#include <cuda_fp16.h>

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

namespace {

const uint32_t kMagic = 0x67676d6c;
const int kQntVersion = 2, kQntVersionFactor = 1000;          // GGML_QNT_VERSION, GGML_QNT_VERSION_FACTOR (ggml.h:212-216)

template <typename T> bool rd(std::ifstream & f, T & v) { f.read(reinterpret_cast<char *>(&v), sizeof(T)); return (bool) f; }
template <typename T> void wr(std::ofstream & f, const T & v) { f.write(reinterpret_cast<const char *>(&v), sizeof(T)); }

// The x*id + c expressions of the reference are ONE fused multiply-add in the pinned build (gcc contracts them under -mfma; oracle/Makefile
// flags), which decides ~1 code in 65 000 differently from the two-rounding form; fmaf() reproduces it.
void put_f16(uint8_t * p, float v) { const __half h = __float2half_rn(v); memcpy(p, &h, 2); }             // GGML_FP32_TO_FP16: round to nearest even

// block_q4_0 = { f16 d; u8 qs[16] }: value = (nibble - 8) * d, low nibbles = elements 0..15, high = 16..31 (ggml-common.h:144-148)
void quantize_row_q4_0(const float * x, uint8_t * y, int64_t k) {
    for (int64_t b = 0; b < k / 32; b++) {
        float amax = 0.0f, mx = 0.0f;
        for (int j = 0; j < 32; j++) { const float v = x[b * 32 + j]; if (amax < fabsf(v)) { amax = fabsf(v); mx = v; } }
        const float d = mx / -8;
        const float id = d ? 1.0f / d : 0.0f;
        uint8_t * blk = y + b * 18;
        put_f16(blk, d);
        for (int j = 0; j < 16; j++) {
            const int i0 = (int8_t) fmaf(x[b * 32 + j], id, 8.5f), i1 = (int8_t) fmaf(x[b * 32 + 16 + j], id, 8.5f);
            const uint8_t q0 = (uint8_t)(i0 < 15 ? i0 : 15), q1 = (uint8_t)(i1 < 15 ? i1 : 15);
            blk[2 + j] = (uint8_t)(q0 | (q1 << 4));
        }
    }
}

// block_q4_1 = { f16 d; f16 m; u8 qs[16] }: value = nibble * d + m (ggml-quants.c:710-745)
void quantize_row_q4_1(const float * x, uint8_t * y, int64_t k) {
    for (int64_t b = 0; b < k / 32; b++) {
        float mn = FLT_MAX, mx = -FLT_MAX;
        for (int j = 0; j < 32; j++) { const float v = x[b * 32 + j]; if (v < mn) mn = v; if (v > mx) mx = v; }
        const float d = (mx - mn) / 15;
        const float id = d ? 1.0f / d : 0.0f;
        uint8_t * blk = y + b * 20;
        put_f16(blk, d); put_f16(blk + 2, mn);
        for (int j = 0; j < 16; j++) {
            const int i0 = (int8_t) fmaf(x[b * 32 + j] - mn, id, 0.5f), i1 = (int8_t) fmaf(x[b * 32 + 16 + j] - mn, id, 0.5f);
            const uint8_t q0 = (uint8_t)(i0 < 15 ? i0 : 15), q1 = (uint8_t)(i1 < 15 ? i1 : 15);
            blk[4 + j] = (uint8_t)(q0 | (q1 << 4));
        }
    }
}

// block_q5_0 = { f16 d; u8 qh[4]; u8 qs[16] }: 5-bit codes, fifth bits in qh (ggml-quants.c:751-793)
void quantize_row_q5_0(const float * x, uint8_t * y, int64_t k) {
    for (int64_t b = 0; b < k / 32; b++) {
        float amax = 0.0f, mx = 0.0f;
        for (int j = 0; j < 32; j++) { const float v = x[b * 32 + j]; if (amax < fabsf(v)) { amax = fabsf(v); mx = v; } }
        const float d = mx / -16;
        const float id = d ? 1.0f / d : 0.0f;
        uint8_t * blk = y + b * 22;
        put_f16(blk, d);
        uint32_t qh = 0;
        for (int j = 0; j < 16; j++) {
            const int i0 = (int8_t) fmaf(x[b * 32 + j], id, 16.5f), i1 = (int8_t) fmaf(x[b * 32 + 16 + j], id, 16.5f);
            const uint8_t q0 = (uint8_t)(i0 < 31 ? i0 : 31), q1 = (uint8_t)(i1 < 31 ? i1 : 31);
            blk[6 + j] = (uint8_t)((q0 & 0x0F) | ((q1 & 0x0F) << 4));
            qh |= ((q0 & 0x10u) >> 4) << (j + 0);
            qh |= ((q1 & 0x10u) >> 4) << (j + 16);
        }
        memcpy(blk + 2, &qh, 4);
    }
}

// block_q5_1 = { f16 d; f16 m; u8 qh[4]; u8 qs[16] } (ggml-quants.c:799-841); no clamp in the reference: (uint8_t)(x + 0.5f)
void quantize_row_q5_1(const float * x, uint8_t * y, int64_t k) {
    for (int64_t b = 0; b < k / 32; b++) {
        float mn = FLT_MAX, mx = -FLT_MAX;
        for (int j = 0; j < 32; j++) { const float v = x[b * 32 + j]; if (v < mn) mn = v; if (v > mx) mx = v; }
        const float d = (mx - mn) / 31;
        const float id = d ? 1.0f / d : 0.0f;
        uint8_t * blk = y + b * 24;
        put_f16(blk, d); put_f16(blk + 2, mn);
        uint32_t qh = 0;
        for (int j = 0; j < 16; j++) {
            const uint8_t q0 = (uint8_t) fmaf(x[b * 32 + j] - mn, id, 0.5f), q1 = (uint8_t) fmaf(x[b * 32 + 16 + j] - mn, id, 0.5f);
            blk[8 + j] = (uint8_t)((q0 & 0x0F) | ((q1 & 0x0F) << 4));
            qh |= ((q0 & 0x10u) >> 4) << (j + 0);
            qh |= ((q1 & 0x10u) >> 4) << (j + 16);
        }
        memcpy(blk + 4, &qh, 4);
    }
}

// block_q8_0 = { f16 d; i8 qs[32] } (ggml-quants.c:848-871): roundf, i.e. halves away from zero
void quantize_row_q8_0(const float * x, uint8_t * y, int64_t k) {
    for (int64_t b = 0; b < k / 32; b++) {
        float amax = 0.0f;
        for (int j = 0; j < 32; j++) { const float v = fabsf(x[b * 32 + j]); amax = amax > v ? amax : v; }
        const float d = amax / 127;
        const float id = d ? 1.0f / d : 0.0f;
        uint8_t * blk = y + b * 34;
        put_f16(blk, d);
        for (int j = 0; j < 32; j++) blk[2 + j] = (uint8_t)(int8_t) roundf(x[b * 32 + j] * id);
    }
}

struct QuantType { int ftype, ttype, block_bytes; void (*row)(const float *, uint8_t *, int64_t); };
// ggml_ftype -> ggml_type (bark.cpp:280-291; ggml.h enum ggml_type: Q4_0 2, Q4_1 3, Q5_0 6, Q5_1 7, Q8_0 8)
const QuantType kQuantTypes[] = {{2, 2, 18, quantize_row_q4_0}, {3, 3, 20, quantize_row_q4_1}, {8, 6, 22, quantize_row_q5_0},
                                 {9, 7, 24, quantize_row_q5_1}, {7, 8, 34, quantize_row_q8_0}};
const QuantType * find_quant(int ftype) { for (const QuantType & q : kQuantTypes) if (q.ftype == ftype) return &q; return nullptr; }

bool ends_with(const std::string & s, const char * suf) { const size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }

// the reference's regex list (bark.cpp:2283-2290), spelled out
bool wants_quantization(const std::string & name) {
    if (name.compare(0, 10, "model/wte/") == 0 || name.compare(0, 14, "model/lm_head/") == 0) return true;
    if (name.compare(0, 7, "model/h") != 0) return false;
    return ends_with(name, "/attn/c_attn/w") || ends_with(name, "/attn/c_proj/w") || ends_with(name, "/mlp/c_fc/w") || ends_with(name, "/mlp/c_proj/w");
}

bool quantize_gpt_section(std::ifstream & fin, std::ofstream & fout, const QuantType & qt, const char * what) {
    const int ftype = qt.ftype;
    int32_t hp[10];
    for (int i = 0; i < 10; i++) if (!rd(fin, hp[i])) { fprintf(stderr, "%s: truncated %s header\n", __func__, what); return false; }
    for (int i = 0; i < 9; i++) wr(fout, hp[i]);
    wr(fout, (int32_t)(kQntVersion * kQntVersionFactor + ftype));
    int32_t n_tensors = 0;
    if (!rd(fin, n_tensors) || n_tensors < 0) return false;
    wr(fout, n_tensors);
    std::vector<char> raw; std::vector<float> f32; std::vector<uint8_t> q;
    for (int t = 0; t < n_tensors; t++) {
        int32_t n_dims = 0, len = 0, ttype = 0, ne[4] = {1, 1, 1, 1};
        if (!rd(fin, n_dims) || !rd(fin, len) || !rd(fin, ttype) || n_dims < 1 || n_dims > 4 || len < 0 || len > 4096) { fprintf(stderr, "%s: malformed tensor record in %s model\n", __func__, what); return false; }
        int64_t nel = 1;
        for (int i = 0; i < n_dims; i++) { if (!rd(fin, ne[i]) || ne[i] <= 0) return false; nel *= ne[i]; }
        std::string name((size_t) len, '\0');
        fin.read(&name[0], len);
        if (!fin) return false;
        const bool quant = wants_quantization(name) && n_dims == 2;
        if (quant) {
            if (ttype != 0 && ttype != 1) { fprintf(stderr, "%s: unsupported ttype %d for integer quantization\n", __func__, ttype); return false; }
            if (ne[0] % 32 != 0) { fprintf(stderr, "%s: tensor '%s': row length %d is not a multiple of 32\n", __func__, name.c_str(), ne[0]); return false; }
            f32.resize((size_t) nel);
            if (ttype == 1) {
                raw.resize((size_t) nel * 2);
                fin.read(raw.data(), (std::streamsize) raw.size());
                const __half * h = reinterpret_cast<const __half *>(raw.data());
                for (int64_t i = 0; i < nel; i++) f32[(size_t) i] = __half2float(h[i]);
            } else {
                fin.read(reinterpret_cast<char *>(f32.data()), (std::streamsize)(nel * 4));
            }
            if (!fin) return false;
            ttype = qt.ttype;
        } else {
            if (ttype != 0 && ttype != 1) { fprintf(stderr, "%s: tensor '%s' has unsupported type %d\n", __func__, name.c_str(), ttype); return false; }
            raw.resize((size_t) nel * (ttype == 0 ? 4 : 2));
            fin.read(raw.data(), (std::streamsize) raw.size());
            if (!fin) return false;
        }
        wr(fout, n_dims); wr(fout, len); wr(fout, ttype);
        for (int i = 0; i < n_dims; i++) wr(fout, ne[i]);
        fout.write(name.data(), len);
        if (quant) {
            q.resize((size_t)(nel / 32 * qt.block_bytes));
            qt.row(f32.data(), q.data(), nel);                                                           // rows are whole numbers of blocks, so one pass over all rows
            fout.write(reinterpret_cast<const char *>(q.data()), (std::streamsize) q.size());
        } else {
            fout.write(raw.data(), (std::streamsize) raw.size());
        }
    }
    return (bool) fout;
}

}  // namespace

extern "C" bool bark_model_quantize(const char * fname_inp, const char * fname_out, enum ggml_ftype ftype) {
    if (!fname_inp || !fname_out) { fprintf(stderr, "%s: null file name\n", __func__); return false; }
    const QuantType * qt = find_quant((int) ftype);
    if (!qt) {
        fprintf(stderr, "%s: ftype %d is not implemented in this build (q4_0, q4_1, q5_0, q5_1, q8_0 are)\n", __func__, (int) ftype);
        return false;
    }
    std::ifstream fin(fname_inp, std::ios::binary);
    if (!fin) { fprintf(stderr, "%s: failed to open '%s' for reading\n", __func__, fname_inp); return false; }
    std::ofstream fout(fname_out, std::ios::binary);
    if (!fout) { fprintf(stderr, "%s: failed to open '%s' for writing\n", __func__, fname_out); return false; }
    uint32_t magic = 0;
    if (!rd(fin, magic) || magic != kMagic) { fprintf(stderr, "%s: invalid model file '%s' (bad magic)\n", __func__, fname_inp); return false; }
    wr(fout, magic);
    uint32_t n_vocab = 0;
    if (!rd(fin, n_vocab)) return false;
    wr(fout, n_vocab);
    std::string word;
    for (uint32_t i = 0; i < n_vocab; i++) {
        uint32_t len = 0;
        if (!rd(fin, len) || len > (1u << 20)) { fprintf(stderr, "%s: failed to read the vocabulary\n", __func__); return false; }
        wr(fout, len);
        word.resize(len);
        if (len) { fin.read(&word[0], len); fout.write(word.data(), len); }
    }
    static const char * names[3] = {"text", "coarse", "fine"};
    for (int s = 0; s < 3; s++)
        if (!quantize_gpt_section(fin, fout, *qt, names[s])) { fprintf(stderr, "%s: failed to quantize %s model\n", __func__, names[s]); return false; }
    if (fin.peek() != std::ifstream::traits_type::eof()) fout << fin.rdbuf();                           // codec section: not quantised, copied verbatim (bark.cpp:2366-2371)
    fout.flush();
    return (bool) fout;
}
