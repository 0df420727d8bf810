// =============================================================================
// xalm_oracle.cpp — CPU restatement of the Xalm reference forward pass.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  Nothing under xalm_b200/ links, imports or calls it.
//
// Parity status: PINNED TO THE REFERENCE ITSELF.  The reference ships no tests / golden vectors, but
// its own C++ forward pass (src/infer.cpp, model.cpp, tensor.cpp, sampler.cpp, tokenizer.cpp) compiles untouched behind
// the NEON / <print> shims of oracle/ref_shim/ (recipe: oracle/Makefile.ref -> oracle/_ref/libxalm_ref.so, dev container
// only).  tests/golden/ref_fwd_*.npz are logits, greedy tokens, sampler probabilities and KV-cache bits produced by that
// library (generator: tests/golden/make_ref_forward.py); tests/test_reference_forward.py requires this restatement to
// reproduce them, and compares it with the live library when it is present.  Also pinned (tests/test_oracle_*.py):
//   * block-quant dequant  == /root/reference/quants.py  (bit-exact, golden vectors)
//   * fp8 decode           == torch float8 casts except the NaN/Inf codes the
//                             reference treats as finite (types.h:302-314)
//   * .xalm layout         == files written by /root/reference/convert.py
//
// Every function cites the reference file:line it restates.  All paths are
// relative to /root/reference/.  Arithmetic is fp32 unless stated, fp16 KV cache,
// exactly as the reference.
// =============================================================================
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef _Float16 f16;

// ---- type ids: src/types.h:505-514 (0..9) and convert.py:56-61 (1007..1012) ----
enum : int {
	T_F32 = 1, T_F16 = 2, T_BF16 = 3, T_F8_E2M5 = 4, T_F8_E3M4 = 5, T_F8_E4M3 = 6, T_F8_E5M2 = 7, T_U8 = 8, T_Q8 = 9,
	T_Q4_0 = 1007, T_Q4_1 = 1008, T_Q5_0 = 1009, T_Q5_1 = 1010, T_Q8_0 = 1011, T_TQ1_0 = 1012,
	T_QI8 = 2007, // convert.py:50 (qi8), decoder convert.py:545-551
};

// (block elements, block bytes): quants.py:45-77; scalar types: types.h:505-514
static bool type_block(int t, int* bs, int* ts) {
	switch (t) {
		case T_F32: *bs = 1; *ts = 4; return true;
		case T_F16: case T_BF16: *bs = 1; *ts = 2; return true;
		case T_F8_E2M5: case T_F8_E3M4: case T_F8_E4M3: case T_F8_E5M2: case T_U8: case T_Q8: case T_QI8:
			*bs = 1; *ts = 1; return true;
		case T_Q4_0: *bs = 32; *ts = 18; return true;
		case T_Q4_1: *bs = 32; *ts = 20; return true;
		case T_Q5_0: *bs = 32; *ts = 22; return true;
		case T_Q5_1: *bs = 32; *ts = 24; return true;
		case T_Q8_0: *bs = 32; *ts = 34; return true;
		case T_TQ1_0: *bs = 256; *ts = 54; return true;
	}
	return false;
}

static inline float u32_as_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float ld_f16(const uint8_t* p) { f16 h; memcpy(&h, p, 2); return (float) h; }

// types.h:302-314 — f8_t<E,M>::to_float: sign to bit 31, low 7 bits to the top of the
// exponent/mantissa field, then multiply by 2^(127 - bias), bias = 2^(E-1) - 1.
// No NaN/Inf special-casing (e4m3 0x7F -> 480, e5m2 0x7C -> 65536).
static inline float f8_to_float(uint8_t b, int E, int M) {
	uint32_t bits = (uint32_t) (b & 0x80) << 24;
	bits |= (uint32_t) (b & 0x7F) << (23 - M);
	const int bias = (1 << (E - 1)) - 1;
	return u32_as_f32(bits) * ldexpf(1.0f, 127 - bias);
}

// types.h:322-325
static inline float bf16_to_f32(uint16_t h) { return u32_as_f32((uint32_t) h << 16); }

// types.h:406-427 — Type::get_float for the scalar types (+ qi8: convert.py:545-551)
static inline float get_float(int t, const void* data, size_t i) {
	const uint8_t* p = (const uint8_t*) data;
	switch (t) {
		case T_F32: { float f; memcpy(&f, p + 4 * i, 4); return f; }
		case T_F16: return ld_f16(p + 2 * i);
		case T_BF16: { uint16_t h; memcpy(&h, p + 2 * i, 2); return bf16_to_f32(h); }
		case T_F8_E2M5: return f8_to_float(p[i], 2, 5);
		case T_F8_E3M4: return f8_to_float(p[i], 3, 4);
		case T_F8_E4M3: return f8_to_float(p[i], 4, 3);
		case T_F8_E5M2: return f8_to_float(p[i], 5, 2);
		case T_Q8: return (1.f / 100.f) * (float) ((const int8_t*) p)[i]; // types.h:423-424
		case T_QI8: return ((float) p[i] / 127.5f) - 1.0f;                // convert.py:550-551
		case T_U8: return (float) p[i];
	}
	return 666.66f; // types.h:426
}

// ---- block formats: one block -> block_size floats -------------------------------
// quants.py:302-313 (Q4_0), :337-350 (Q4_1), :376-393 (Q5_0), :419-438 (Q5_1),
// :458-464 (Q8_0), :664-683 (TQ1_0).  Layouts: SURVEY.md Appendix B.
static void dequant_block(int t, const uint8_t* b, float* y) {
	switch (t) {
		case T_Q8_0: {
			const float d = ld_f16(b);
			const int8_t* q = (const int8_t*) (b + 2);
			for (int j = 0; j < 32; j++) y[j] = (float) q[j] * d;
			return;
		}
		case T_Q4_0: {
			const float d = ld_f16(b);
			const uint8_t* qs = b + 2;
			for (int j = 0; j < 16; j++) {
				y[j] = d * (float) ((int) (qs[j] & 0x0F) - 8);
				y[j + 16] = d * (float) ((int) (qs[j] >> 4) - 8);
			}
			return;
		}
		case T_Q4_1: {
			const float d = ld_f16(b), m = ld_f16(b + 2);
			const uint8_t* qs = b + 4;
			for (int j = 0; j < 16; j++) {
				// (d * q) + m as two separately rounded fp32 ops (numpy); d*q is exact
				// (11-bit x 4-bit), so a fused multiply-add gives the same bits.
				volatile float lo = d * (float) (qs[j] & 0x0F);
				volatile float hi = d * (float) (qs[j] >> 4);
				y[j] = lo + m;
				y[j + 16] = hi + m;
			}
			return;
		}
		case T_Q5_0: {
			const float d = ld_f16(b);
			uint32_t qh; memcpy(&qh, b + 2, 4);
			const uint8_t* qs = b + 6;
			for (int j = 0; j < 16; j++) {
				const int lo = (qs[j] & 0x0F) | (((qh >> j) & 1) << 4);
				const int hi = (qs[j] >> 4) | (((qh >> (j + 16)) & 1) << 4);
				y[j] = d * (float) (lo - 16);
				y[j + 16] = d * (float) (hi - 16);
			}
			return;
		}
		case T_Q5_1: {
			const float d = ld_f16(b), m = ld_f16(b + 2);
			uint32_t qh; memcpy(&qh, b + 4, 4);
			const uint8_t* qs = b + 8;
			for (int j = 0; j < 16; j++) {
				const int lo = (qs[j] & 0x0F) | (((qh >> j) & 1) << 4);
				const int hi = (qs[j] >> 4) | (((qh >> (j + 16)) & 1) << 4);
				volatile float plo = d * (float) lo;
				volatile float phi = d * (float) hi;
				y[j] = plo + m;
				y[j + 16] = phi + m;
			}
			return;
		}
		case T_TQ1_0: {
			// [0:32] qs0 (5 trits/byte) [32:48] qs1 (5 trits/byte) [48:52] qh (4 trits/byte) [52:54] d
			static const uint8_t pow3[5] = {1, 3, 9, 27, 81};
			const float d = ld_f16(b + 52);
			int o = 0;
			for (int k = 0; k < 5; k++)
				for (int B = 0; B < 32; B++) {
					const uint8_t q = (uint8_t) (b[B] * pow3[k]);
					y[o++] = d * (float) ((int) (((uint16_t) q * 3) >> 8) - 1);
				}
			for (int k = 0; k < 5; k++)
				for (int B = 0; B < 16; B++) {
					const uint8_t q = (uint8_t) (b[32 + B] * pow3[k]);
					y[o++] = d * (float) ((int) (((uint16_t) q * 3) >> 8) - 1);
				}
			for (int k = 0; k < 4; k++)
				for (int B = 0; B < 4; B++) {
					const uint8_t q = (uint8_t) (b[48 + B] * pow3[k]);
					y[o++] = d * (float) ((int) (((uint16_t) q * 3) >> 8) - 1);
				}
			return;
		}
	}
}

extern "C" {

// bytes needed for n_elems elements of type t (n_elems % block == 0), or -1
long long orc_type_nbytes(int t, long long n_elems) {
	int bs, ts;
	if (!type_block(t, &bs, &ts) || n_elems % bs) return -1;
	return n_elems / bs * ts;
}

// Dequantise n_elems consecutive elements (row-major) to fp32.
int orc_dequant(int t, const void* src, long long n_elems, float* dst) {
	int bs, ts;
	if (!type_block(t, &bs, &ts) || n_elems % bs) return 1;
	if (bs == 1) {
		for (long long i = 0; i < n_elems; i++) dst[i] = get_float(t, src, (size_t) i);
		return 0;
	}
	const uint8_t* p = (const uint8_t*) src;
	const long long nb = n_elems / bs;
#pragma omp parallel for schedule(static)
	for (long long b = 0; b < nb; b++) dequant_block(t, p + b * ts, dst + b * bs);
	return 0;
}

// ---- matmul: infer.cpp:104-135 (template) and :185-216 (dispatch) ----------------
// W (d,n) row-major @ x (n,) -> xout (d,).  acc_mode: 0 = strict sequential fp32 (the
// loop as written), 1 = fp32 with `omp simd`-style reassociation (8 partial sums; what
// a vectorising compiler makes of infer.cpp:121), 2 = fp64 accumulate (noise bound).
static void row_dequant(int t, const uint8_t* wrow, int n, float* tmp) {
	int bs, ts;
	type_block(t, &bs, &ts);
	if (bs == 1) { for (int j = 0; j < n; j++) tmp[j] = get_float(t, wrow, (size_t) j); return; }
	for (int b = 0; b < n / bs; b++) dequant_block(t, wrow + (size_t) b * ts, tmp + b * bs);
}

static inline float dot_mode(const float* w, const float* x, int n, int acc_mode) {
	if (acc_mode == 0) {
		float val = 0.0f;
		for (int j = 0; j < n; j++) { volatile float p = w[j] * x[j]; val += p; } // no fma contraction, strict order
		return val;
	}
	if (acc_mode == 2) {
		double val = 0.0;
		for (int j = 0; j < n; j++) val += (double) w[j] * (double) x[j];
		return (float) val;
	}
	float val = 0.0f;
#pragma omp simd reduction(+ : val)
	for (int j = 0; j < n; j++) val += w[j] * x[j];
	return val;
}

int orc_matmul(float* xout, const float* x, const void* w, int t, int n, int d, int acc_mode) {
	int bs, ts;
	if (!type_block(t, &bs, &ts) || t == T_U8) { fprintf(stderr, "matmul: unsupported data type: %d\n", t); return 1; } // infer.cpp:211-214
	if (n % 32 || d % 32) return 2; // infer.cpp:110-111 asserts
	if (n % bs) return 2;
	const size_t row_bytes = (size_t) n / bs * ts;
	const uint8_t* wb = (const uint8_t*) w;
#pragma omp parallel
	{
		std::vector<float> tmp((size_t) n);
#pragma omp for schedule(static)
		for (int i = 0; i < d; i++) {
			const uint8_t* wrow = wb + (size_t) i * row_bytes;
			if (t == T_F32) {
				xout[i] = dot_mode((const float*) wrow, x, n, acc_mode);
			} else {
				row_dequant(t, wrow, n, tmp.data());
				xout[i] = dot_mode(tmp.data(), x, n, acc_mode);
			}
		}
	}
	return 0;
}

// ---- rmsnorm: infer.cpp:224-251 (weight F32 or BF16 only) --------------------------
int orc_rmsnorm(float* o, const float* x, const void* weight, int wtype, int size, float eps) {
	if (wtype != T_F32 && wtype != T_BF16) return 1; // infer.cpp:248-249 throws
	float rms = 0.0f;
	for (int i = 0; i < size; ++i) { volatile float p = x[i] * x[i]; rms += p; }
	rms = sqrtf(rms / (float) size + eps);
	const float scale = 1.0f / rms;
	for (int i = 0; i < size; ++i) {
		volatile float xs = x[i] * scale;
		o[i] = xs * get_float(wtype, weight, (size_t) i);
	}
	return 0;
}

// ---- softmax: infer.cpp:280-297 -----------------------------------------------------
void orc_softmax(float* o, const float* x, int size) {
	float score_max = std::numeric_limits<float>::lowest();
	for (int i = 0; i < size; ++i) if (x[i] > score_max) score_max = x[i];
	float score_sum = 0.0f;
	for (int i = 0; i < size; ++i) { o[i] = expf(x[i] - score_max); score_sum += o[i]; }
	for (int i = 0; i < size; ++i) o[i] /= score_sum;
}

// infer.cpp:299-303
float orc_gelu(float x) { return 0.5f * x * (1.0f + tanhf(0.797885f * (x + 0.044715f * x * x * x))); }
float orc_silu(float x) { return x / (1.0f + expf(-x)); }
float orc_clip(float x, float v) { return x < -v ? -v : (x > v ? v : x); }

// ---- rope: infer.cpp:305-322 ----------------------------------------------------------
void orc_rope(float* vec, int d, int head_dim, int pos, float theta, int rotary_dim) {
	for (int i = 0; i < d; i += 2) {
		const int j_head = i % head_dim;
		const float freq = j_head >= rotary_dim ? 0.f : 1.0f / powf(theta, (float) j_head / (float) rotary_dim);
		const float val = pos * freq;
		const float fcr = cosf(val);
		const float fci = sinf(val);
		const float v0 = vec[i];
		const float v1 = vec[i + 1];
		volatile float a = v0 * fcr, b = v1 * fci, c = v0 * fci, e = v1 * fcr;
		vec[i] = a - b;
		vec[i + 1] = c + e;
	}
}

// ---- attn: infer.cpp:325-359 -------------------------------------------------------------
void orc_attn(float* xout, float* atth, const float* qh, const uint16_t* kh_, const uint16_t* vh_, int head_dim,
              int n_kv_heads, int kv_len) {
	const f16* kh = (const f16*) kh_;
	const f16* vh = (const f16*) vh_;
	const int kv_stride = n_kv_heads * head_dim;
	const float sqrt_head_dim = 1.0f / sqrtf((float) head_dim);
	for (int t = 0; t < kv_len; ++t) {
		float score = 0.0f;
		for (int i = 0; i < head_dim; ++i) score += qh[i] * (float) kh[(size_t) t * kv_stride + i];
		atth[t] = score * sqrt_head_dim;
	}
	orc_softmax(atth, atth, kv_len);
	for (int i = 0; i < head_dim; ++i) {
		float vi = 0.0f;
		for (int t = 0; t < kv_len; ++t) vi += atth[t] * (float) vh[(size_t) t * kv_stride + i];
		xout[i] = vi;
	}
}

// ---- mha_cpu: infer.cpp:498-517 ------------------------------------------------------------
void orc_mha(float* xout, float* att, const uint16_t* kb, const uint16_t* vb, const float* q, int head_dim, int kv_len,
             int max_seq_len, int n_heads, int n_kv_heads) {
	const int q_per_kv_head = n_heads / n_kv_heads;
#pragma omp parallel for
	for (int h = 0; h < n_heads; h++) {
		const int kv_head_offset = (h / q_per_kv_head) * head_dim;
		orc_attn(xout + (size_t) head_dim * h, att + (size_t) max_seq_len * h, q + (size_t) head_dim * h, kb + kv_head_offset,
		         vb + kv_head_offset, head_dim, n_kv_heads, kv_len);
	}
}

// ---- Config (model.h:25-42), Model/Block/InferenceState (model.h:96-284) ----------------------
struct orc_config {
	int dim, hidden_dim, head_dim, n_layers, n_heads, n_kv_heads, vocab_size, max_seq_len;
	float rope_theta;
	int rotary_dim;
	float norm_eps;
	int act;       // 0 = GELU, 1 = SILU (model.h:12-15)
	int norm_type; // 0 = RMSNorm
	float qkv_clip;
	int tie_word_embeddings;
};

struct OTensor { const void* data = nullptr; int type = 0; };

struct OBlock {
	OTensor rms_att, rms_ffn, wq, wk, wv, wo, w1, w2, w3;
	std::vector<uint16_t> key_cache, value_cache; // fp16 bits, (max_seq_len, kv_dim) model.h:222-223
};

struct orc_model {
	orc_config c;
	OTensor embed, rms_final, wcls;
	std::vector<OBlock> blocks;
	// InferenceState model.h:96-156
	std::vector<float> x, xb, xb2, hb, hb2, q, k, v, att, logits;
	int acc_mode = 1;
};

orc_model* orc_model_create(const orc_config* c) {
	orc_model* m = new orc_model();
	m->c = *c;
	m->blocks.resize(c->n_layers);
	const size_t kv_dim = (size_t) c->n_kv_heads * c->head_dim;
	for (auto& b : m->blocks) {
		// the reference leaves these uninitialised (model.cpp:102-103); zero is as good as any
		b.key_cache.assign((size_t) c->max_seq_len * kv_dim, 0);
		b.value_cache.assign((size_t) c->max_seq_len * kv_dim, 0);
	}
	m->x.assign(c->dim, 0); m->xb.assign(c->dim, 0); m->xb2.assign(c->dim, 0);
	m->hb.assign(c->hidden_dim, 0); m->hb2.assign(c->hidden_dim, 0);
	m->q.assign((size_t) c->n_heads * c->head_dim, 0);
	m->k.assign(kv_dim, 0); m->v.assign(kv_dim, 0);
	m->att.assign((size_t) c->n_heads * c->max_seq_len, 0);
	m->logits.assign(c->vocab_size, 0);
	return m;
}

void orc_model_destroy(orc_model* m) { delete m; }
void orc_model_set_acc_mode(orc_model* m, int mode) { m->acc_mode = mode; }

// Tensor names: model.cpp:83-114.  The pointer is borrowed (caller keeps it alive).
int orc_model_set_tensor(orc_model* m, const char* name, int type, const void* data) {
	const std::string n(name);
	OTensor t{data, type};
	if (n == "embed.weight") { m->embed = t; if (m->c.tie_word_embeddings) m->wcls = t; return 0; } // model.cpp:112-114
	if (n == "output.norm.weight") { m->rms_final = t; return 0; }
	if (n == "output.weight") { if (!m->c.tie_word_embeddings) m->wcls = t; return 0; }
	int l = -1; char rest[64] = {0};
	if (sscanf(name, "l.%d.%63s", &l, rest) == 2 && l >= 0 && l < m->c.n_layers) {
		OBlock& b = m->blocks[l];
		const std::string r(rest);
		if (r == "attn.norm.weight") b.rms_att = t;
		else if (r == "mlp.norm.weight") b.rms_ffn = t;
		else if (r == "attn.q.weight") b.wq = t;
		else if (r == "attn.k.weight") b.wk = t;
		else if (r == "attn.v.weight") b.wv = t;
		else if (r == "attn.down.weight") b.wo = t;
		else if (r == "mlp.gate.weight") b.w1 = t;
		else if (r == "mlp.down.weight") b.w2 = t;
		else if (r == "mlp.up.weight") b.w3 = t;
		else return 1;
		return 0;
	}
	return 1;
}

// Block::_block_cpu — infer.cpp:365-496
static int block_forward(orc_model* m, OBlock& b, int pos, int kv_sink, int kv_pos, int kv_len) {
	const orc_config& c = m->c;
	const int am = m->acc_mode;
	if (orc_rmsnorm(m->xb.data(), m->x.data(), b.rms_att.data, b.rms_att.type, c.dim, c.norm_eps)) return 1;
	const int q_dim = c.n_heads * c.head_dim;
	const int kv_dim = c.n_kv_heads * c.head_dim;
	if (orc_matmul(m->q.data(), m->xb.data(), b.wq.data, b.wq.type, c.dim, q_dim, am)) return 2;
	if (orc_matmul(m->k.data(), m->xb.data(), b.wk.data, b.wk.type, c.dim, kv_dim, am)) return 2;
	if (orc_matmul(m->v.data(), m->xb.data(), b.wv.data, b.wv.type, c.dim, kv_dim, am)) return 2;
	for (int i = 0; i < q_dim; ++i) m->q[i] = orc_clip(m->q[i], c.qkv_clip);
	for (int i = 0; i < kv_dim; ++i) { m->k[i] = orc_clip(m->k[i], c.qkv_clip); m->v[i] = orc_clip(m->v[i], c.qkv_clip); }
	f16* kb = (f16*) b.key_cache.data();
	f16* vb = (f16*) b.value_cache.data();
	orc_rope(m->q.data(), q_dim, c.head_dim, pos, c.rope_theta, c.rotary_dim);
	orc_rope(m->k.data(), kv_dim, c.head_dim, pos, c.rope_theta, c.rotary_dim);
	for (int i = 0; i < kv_dim; ++i) { // infer.cpp:411-414 (fp32 -> fp16 RNE)
		kb[(size_t) kv_pos * kv_dim + i] = (f16) m->k[i];
		vb[(size_t) kv_pos * kv_dim + i] = (f16) m->v[i];
	}
	for (int r = 0; r < kv_sink; r++) { // infer.cpp:421-431
		for (int i = 0; i < kv_dim; ++i) m->k[i] = (float) kb[(size_t) r * kv_dim + i];
		orc_rope(m->k.data(), kv_dim, c.head_dim, 1, c.rope_theta, c.rotary_dim);
		for (int i = 0; i < kv_dim; i++) kb[(size_t) r * kv_dim + i] = (f16) m->k[i];
	}
	orc_mha(m->xb2.data(), m->att.data(), b.key_cache.data(), b.value_cache.data(), m->q.data(), c.head_dim, kv_len,
	        c.max_seq_len, c.n_heads, c.n_kv_heads); // infer.cpp:434-444
	if (orc_matmul(m->hb.data(), m->xb2.data(), b.wo.data, b.wo.type, q_dim, c.dim, am)) return 2;
	for (int i = 0; i < c.dim; ++i) m->x[i] += m->hb[i];
	if (orc_rmsnorm(m->xb.data(), m->x.data(), b.rms_ffn.data, b.rms_ffn.type, c.dim, c.norm_eps)) return 1;
	if (orc_matmul(m->hb.data(), m->xb.data(), b.w1.data, b.w1.type, c.dim, c.hidden_dim, am)) return 2;
	if (orc_matmul(m->hb2.data(), m->xb.data(), b.w3.data, b.w3.type, c.dim, c.hidden_dim, am)) return 2;
	if (c.act == 0) for (int i = 0; i < c.hidden_dim; ++i) m->hb[i] = orc_gelu(m->hb[i]) * m->hb2[i];
	else for (int i = 0; i < c.hidden_dim; ++i) m->hb[i] = orc_silu(m->hb[i]) * m->hb2[i];
	if (orc_matmul(m->xb2.data(), m->hb.data(), b.w2.data, b.w2.type, c.hidden_dim, c.dim, am)) return 2;
	for (int i = 0; i < c.dim; ++i) m->x[i] += m->xb2[i];
	return 0;
}

// Model::_forward_cpu — infer.cpp:604-638 ; _copy_embedding — infer.cpp:553-602
// mode: 0 = HYDRATE_KV_CACHE, 1 = OUTPUT_LOGITS (model.h:249-252)
int orc_forward(orc_model* m, int token, int pos, int mode) {
	const orc_config& c = m->c;
	{
		int bs, ts;
		if (!type_block(m->embed.type, &bs, &ts)) return 3;
		if (bs == 1) {
			for (int i = 0; i < c.dim; ++i) m->x[i] = get_float(m->embed.type, m->embed.data, (size_t) token * c.dim + i);
		} else { // block formats: extension (the reference cannot load them)
			const uint8_t* row = (const uint8_t*) m->embed.data + (size_t) token * (c.dim / bs * ts);
			row_dequant(m->embed.type, row, c.dim, m->x.data());
		}
	}
	const int KV_SINKS = 2; // model.h:10
	const int kv_sink = pos >= c.max_seq_len ? KV_SINKS : 0;
	const int kv_pos = kv_sink + (pos - kv_sink) % (c.max_seq_len - kv_sink);
	const int kv_len = pos >= c.max_seq_len ? c.max_seq_len : pos + 1;
	for (auto& b : m->blocks) {
		const int rc = block_forward(m, b, pos, kv_sink, kv_pos, kv_len);
		if (rc) return rc;
	}
	if (mode == 0) return 0;
	if (orc_rmsnorm(m->x.data(), m->x.data(), m->rms_final.data, m->rms_final.type, c.dim, c.norm_eps)) return 1;
	if (orc_matmul(m->logits.data(), m->x.data(), m->wcls.data, m->wcls.type, c.dim, c.vocab_size, m->acc_mode)) return 2;
	return 0;
}

float* orc_logits(orc_model* m) { return m->logits.data(); }
float* orc_state(orc_model* m, int which) {
	switch (which) {
		case 0: return m->x.data(); case 1: return m->xb.data(); case 2: return m->xb2.data();
		case 3: return m->hb.data(); case 4: return m->hb2.data(); case 5: return m->q.data();
		case 6: return m->k.data(); case 7: return m->v.data(); case 8: return m->att.data();
		case 9: return m->logits.data();
	}
	return nullptr;
}
uint16_t* orc_kv(orc_model* m, int layer, int which) {
	return which == 0 ? m->blocks[layer].key_cache.data() : m->blocks[layer].value_cache.data();
}

// ---- Sampler: sampler.cpp:3-30 — note the FLT_MIN (not lowest) seed, SURVEY.md §0.8 -----------
float orc_sample_prob(const float* logits, int vocab_size, int index) {
	float max_val = std::numeric_limits<float>::min();
	for (int i = 0; i < vocab_size; ++i) if (logits[i] > max_val) max_val = logits[i];
	float sum = 0;
	for (int i = 0; i < vocab_size; ++i) sum += expf(logits[i] - max_val);
	return expf(logits[index] - max_val) / sum;
}
int orc_sample_argmax(const float* logits, int vocab_size) {
	int argmax = 0;
	float max_val = std::numeric_limits<float>::min();
	for (int i = 0; i < vocab_size; ++i) if (logits[i] > max_val) { max_val = logits[i]; argmax = i; }
	return argmax;
}

// ---- Model::active_bytes — model.cpp:12-35, 64-bit, block formats via ts/bs -------------------
long long orc_active_bytes(orc_model* m, long long pos) {
	const orc_config& c = m->c;
	auto wb = [](const OTensor& t, long long elems) -> long long {
		int bs = 1, ts = 1; type_block(t.type, &bs, &ts); return elems / bs * ts;
	};
	long long bytes = 0;
	bytes += wb(m->embed, c.dim);
	bytes += wb(m->rms_final, c.dim);
	bytes += wb(m->wcls, (long long) c.vocab_size * c.dim);
	const long long q_dim = (long long) c.n_heads * c.head_dim, kv_dim = (long long) c.n_kv_heads * c.head_dim;
	for (auto& b : m->blocks) {
		bytes += wb(b.rms_att, c.dim) + wb(b.rms_ffn, c.dim);
		bytes += wb(b.wq, q_dim * c.dim) + wb(b.wk, kv_dim * c.dim) + wb(b.wv, kv_dim * c.dim) + wb(b.wo, q_dim * c.dim);
		bytes += wb(b.w1, (long long) c.dim * c.hidden_dim) + wb(b.w2, (long long) c.dim * c.hidden_dim) +
		         wb(b.w3, (long long) c.dim * c.hidden_dim);
		const long long kv_len = std::min<long long>(c.max_seq_len, pos + 1);
		bytes += 2 * kv_len * kv_dim * 2;
	}
	return bytes;
}

// torchrun exports OMP_NUM_THREADS=1 into every rank; the CPU-baseline leg asks for all cores explicitly
void orc_set_threads(int n) {
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#else
	(void) n;
#endif
}

int orc_num_threads() {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

} // extern "C"
