// bark_model_quantize (bark.h:229-232): f32 / f16 ggml_weights.bin -> q4_0 ggml_weights.bin, host only.
//
// Same container walk as the reference (bark_model_quantize bark.cpp:2300-2377, bark_model_weights_quantize
// bark.cpp:2234-2298, ggml_quantize_weights bark.cpp:272-470): magic and vocabulary copied, each GPT section re-emitted with
// ftype = GGML_QNT_VERSION * 1000 + ftype, the 2-D tensors named wte / lm_head / c_attn / c_proj / c_fc / mlp c_proj
// quantised row by row with quantize_row_q4_0_ref (ggml-quants.c:668-703), everything else and the whole codec section copied
// byte for byte.  Output files are byte-identical to the reference tool's (tests/test_quantize.py).  Only q4_0 is
// implemented (the GPU path reads f32, f16 and q4_0); other ftypes are rejected with a message.
#include "../../include/bark.h"

#include <cuda_fp16.h>

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

// block_q4_0 = { f16 d; u8 qs[16] }: value = (nibble - 8) * d, low nibbles = elements 0..15, high = 16..31 (ggml-common.h:144-148)
void quantize_row_q4_0(const float * x, uint8_t * y, int64_t k) {
    for (int64_t b = 0; b < k / 32; b++) {
        float amax = 0.0f, mx = 0.0f;
        for (int j = 0; j < 32; j++) { const float v = x[b * 32 + j]; if (amax < fabsf(v)) { amax = fabsf(v); mx = v; } }
        const float d = mx / -8;
        const float id = d ? 1.0f / d : 0.0f;
        const __half dh = __float2half_rn(d);                                                          // GGML_FP32_TO_FP16: round to nearest even
        uint8_t * blk = y + b * 18;
        memcpy(blk, &dh, 2);
        for (int j = 0; j < 16; j++) {
            // x*id + 8.5f is ONE fused multiply-add in the pinned reference build (gcc contracts it under -mfma; oracle/Makefile
            // flags), which decides 1 nibble in ~65 000 differently from the two-rounding form
            const int i0 = (int8_t) fmaf(x[b * 32 + j], id, 8.5f), i1 = (int8_t) fmaf(x[b * 32 + 16 + j], id, 8.5f);
            const uint8_t q0 = (uint8_t)(i0 < 15 ? i0 : 15), q1 = (uint8_t)(i1 < 15 ? i1 : 15);
            blk[2 + j] = (uint8_t)(q0 | (q1 << 4));
        }
    }
}

bool ends_with(const std::string & s, const char * suf) { const size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }

// the reference's regex list (bark.cpp:2283-2290), spelled out
bool wants_quantization(const std::string & name) {
    if (name.compare(0, 10, "model/wte/") == 0 || name.compare(0, 14, "model/lm_head/") == 0) return true;
    if (name.compare(0, 7, "model/h") != 0) return false;
    return ends_with(name, "/attn/c_attn/w") || ends_with(name, "/attn/c_proj/w") || ends_with(name, "/mlp/c_fc/w") || ends_with(name, "/mlp/c_proj/w");
}

bool quantize_gpt_section(std::ifstream & fin, std::ofstream & fout, int ftype, const char * what) {
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
            ttype = 2;                                                                                  // GGML_TYPE_Q4_0
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
            q.resize((size_t)(nel / 32 * 18));
            quantize_row_q4_0(f32.data(), q.data(), nel);                                                // rows are whole numbers of blocks, so one pass over all rows
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
    if ((int) ftype != 2) {                                                                             // GGML_FTYPE_MOSTLY_Q4_0
        fprintf(stderr, "%s: only q4_0 (ftype 2) is implemented in this build (got ftype %d)\n", __func__, (int) ftype);
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
        if (!quantize_gpt_section(fin, fout, (int) ftype, names[s])) { fprintf(stderr, "%s: failed to quantize %s model\n", __func__, names[s]); return false; }
    if (fin.peek() != std::ifstream::traits_type::eof()) fout << fin.rdbuf();                           // codec section: not quantised, copied verbatim (bark.cpp:2366-2371)
    fout.flush();
    return (bool) fout;
}
