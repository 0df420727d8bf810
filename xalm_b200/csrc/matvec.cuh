// matvec.cuh — the decode hot kernel: xout = W(d,n) · x(n), W in any reference weight format, fp32 activations and
// accumulation (infer.cpp:104-135), with the surrounding element-wise ops of Block::_block_cpu fused in:
//
//   prologue (optional)   x' = rmsnorm(x) * g                                   infer.cpp:224-236, :376, :457, :628
//   epilogue STORE        out[i]  = y_i                                         classifier, infer.cpp:637
//            RESIDUAL     out[i] += y_i                                         infer.cpp:447-452, :490-494
//            GLU          out[i]  = act(y_i^{W1}) * y_i^{W3}                    infer.cpp:468-488
//            QKV          clip -> RoPE(q,k) -> q fp32, k/v fp16 into the KV ring at kv_pos; the block that owns
//                         virtual row 0 also re-rotates the attention sinks     infer.cpp:388-431
//
// Mapping.  HBM-bound: every weight byte is read exactly once, with 128-bit streaming loads that bypass L1
// allocation; x is small (<= 114 KB) and is read through L1, where it stays resident for all warps of the SM.
// One warp owns R consecutive rows so each x chunk it loads is reused R times from registers (otherwise the
// 4-byte activations would cost more L1 bandwidth than the <= 1-byte weights cost HBM bandwidth).  KS warps of a
// CTA split K for the same rows when rows are scarce (Wo/W2: 4096 rows must still cover 148 SMs evenly); partial
// sums are combined in a fixed order through shared memory, so results are deterministic run to run.
//
// PDL.  The kernel is launched with programmatic stream serialisation: it fetches its first weight fragments —
// which no earlier kernel writes — BEFORE griddepcontrol.wait, so the HBM stream does not drain at kernel
// boundaries; activations are only touched after the wait.
#pragma once
#include <math_constants.h>

#include "formats.cuh"

namespace xalm {

enum { EPI_STORE = 0, EPI_RESIDUAL = 1, EPI_GLU = 2, EPI_QKV = 3 };

struct MatvecArgs {
	WMat w;              // weights; for GLU rows [0,d) are W1 and [glu_off, glu_off+d) are W3
	const float* x;      // (n,) input activations
	int n;               // K
	int d;               // number of OUTPUT values (GLU: hidden_dim; else rows)
	int epi;
	int glu_off;         // row offset of W3 inside w (GLU only)
	int act;             // xalm_act (GLU only)
	// rmsnorm prologue
	const uint8_t* norm_w; // nullptr = no norm
	int norm_type;       // XALM_F32 or XALM_BF16
	float norm_eps;
	// outputs
	float* out;          // STORE / RESIDUAL / GLU target, QKV: q (q_dim,)
	// QKV epilogue
	const StepParams* step;
	__half* k_cache;     // (max_seq_len, kv_dim)
	__half* v_cache;
	const float* rope_freq; // (head_dim/2,) 1/powf(theta, j/rotary_dim) computed on the host (0 beyond rotary_dim)
	int q_dim, kv_dim, head_dim;
	float qkv_clip;
	// one-byte LUT formats
	int lut_type;        // 0 = none, else the type id whose 256 values fill the shared-memory table
	// tail prefetch: when this CTA's producer has issued its last weight load it pulls its share of the NEXT kernel's first
	// bytes into L2, so that kernel's ramp-up reads hit L2 instead of paying an HBM round trip after the dependency wait
	const uint8_t* pf_ptr;       // next kernel's weights (or nullptr)
	unsigned long long pf_bytes;
	int pf_kv;                   // 1: also prefetch rows [0, kv_len) of k_cache / v_cache (the attention kernel that follows QKV)
	// rmsnorm weights of the NEXT norm-fused kernel (read once per token, i.e. from DRAM): CTA 0 pulls them into L2 at entry.  That
	// kernel requests them before its dependency wait; coming from DRAM behind its own weight stream they took ~2.5 us, and the
	// activation loads issued after them return in order behind them (profiles/r2_decode_timeline.md)
	const uint8_t* pf_norm_ptr;
	unsigned int pf_norm_bytes;
	// tensor-parallel exchange fused into the matvec kernels (matvec_tma.cuh), LL style: the row-split matvecs (Wo, W2) PUSH every
	// partial row as an 8-byte {value, sequence tag} word straight into every rank's receive slot over NVLink (posted stores, no
	// fence, no flag, no extra kernel); the next kernel's rmsnorm prologue polls the tags of the words it needs and adds the
	// partials (fixed rank order) to the residual stream.
	uint2* push_dst[8];          // per destination rank: its receive slot for THIS rank's partial, (dim,) {float bits, tag}
	int n_push;                  // 0 = off
	int push_idx;                // exchange index inside the token: tag = step->ar_base + push_idx + 1
	const uint2* recv;           // local receive slot: (n_recv, dim) words
	int n_recv;                  // 0 = off
	int recv_idx;
	float* x_out;                // the first CTAs store x + sum(partials) here: the residual stream after the exchange
	unsigned int* err_flag;      // set to 1 when a wait for a peer's words times out (the host turns it into XALM_ERR_COMM)
	uint2* xl;                   // local (dim,) {value, tag} words: the reducing CTAs publish the summed stream here, every CTA polls it
	// L2 prefetcher hand-shake (prefetch.cuh): block 0 publishes "kernel #prog_idx of this token has started"
	unsigned int* progress;
	int prog_idx;
};

__device__ __forceinline__ float act_gelu(float x) { return 0.5f * x * (1.0f + tanhf(0.797885f * (x + 0.044715f * x * x * x))); }
__device__ __forceinline__ float act_silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float clipf(float x, float v) { return x < -v ? -v : (x > v ? v : x); }

// rotate the pair (v0, v1) that sits at even head offset j_head by pos * freq (infer.cpp:305-322)
__device__ __forceinline__ void rope_pair(float& v0, float& v1, int j_head, int pos, const float* freq_tab) {
	const float freq = freq_tab[j_head >> 1];
	const float val = (float) pos * freq;
	float fci, fcr;
	sincosf(val, &fci, &fcr);
	const float a = v0, b = v1;
	v0 = a * fcr - b * fci;
	v1 = a * fci + b * fcr;
}

// Re-rotate the kv_sink attention-sink keys by one position, through fp16 (infer.cpp:421-431).
__device__ inline void rotate_sinks(const MatvecArgs& a, int kv_sink) {
	const int pairs = a.kv_dim / 2;
	for (int i = threadIdx.x; i < kv_sink * pairs; i += blockDim.x) {
		const int r = i / pairs, p = i % pairs;
		__half2* kp = reinterpret_cast<__half2*>(a.k_cache + (size_t) r * a.kv_dim) + p;
		float2 v = __half22float2(*kp);
		rope_pair(v.x, v.y, (2 * p) % a.head_dim, 1, a.rope_freq);
		*kp = __floats2half2_rn(v.x, v.y);
	}
}

template <int R>
__device__ __forceinline__ void epilogue(const MatvecArgs& a, int row0, const float (&y)[R]) {
	// called by ONE thread per row group with the R finished dot products; row0 = first virtual row
	if (a.epi == EPI_STORE) {
#pragma unroll
		for (int r = 0; r < R; r++)
			if (row0 + r < a.d) a.out[row0 + r] = y[r];
	} else if (a.epi == EPI_RESIDUAL) {
#pragma unroll
		for (int r = 0; r < R; r++)
			if (row0 + r < a.d) a.out[row0 + r] += y[r];
	} else if (a.epi == EPI_GLU) {
		// rows [0,R/2) came from W1, [R/2,R) from W3, output index = row0/2 + r
		const int o0 = row0 / 2;
#pragma unroll
		for (int r = 0; r < R / 2; r++) {
			if (o0 + r < a.d) {
				const float g = a.act == XALM_SILU ? act_silu(y[r]) : act_gelu(y[r]);
				a.out[o0 + r] = g * y[R / 2 + r];
			}
		}
	} else { // EPI_QKV
		const int pos = a.step->pos, kv_pos = a.step->kv_pos;
#pragma unroll
		for (int r = 0; r < R; r += 2) {
			const int row = row0 + r;
			if (row >= a.d) break;
			float v0 = clipf(y[r], a.qkv_clip), v1 = clipf(y[r + 1], a.qkv_clip);
			if (row < a.q_dim) {
				rope_pair(v0, v1, row % a.head_dim, pos, a.rope_freq);
				a.out[row] = v0;
				a.out[row + 1] = v1;
			} else if (row < a.q_dim + a.kv_dim) {
				const int i = row - a.q_dim;
				rope_pair(v0, v1, i % a.head_dim, pos, a.rope_freq);
				*reinterpret_cast<__half2*>(a.k_cache + (size_t) kv_pos * a.kv_dim + i) = __floats2half2_rn(v0, v1);
			} else {
				const int i = row - a.q_dim - a.kv_dim;
				*reinterpret_cast<__half2*>(a.v_cache + (size_t) kv_pos * a.kv_dim + i) = __floats2half2_rn(v0, v1);
			}
		}
	}
}

// virtual row (what the epilogue sees) -> physical weight row
__device__ __forceinline__ int phys_row(const MatvecArgs& a, int row0, int r, int R) {
	if (a.epi == EPI_GLU) {
		const int o0 = row0 / 2;
		return r < R / 2 ? o0 + r : a.glu_off + o0 + (r - R / 2);
	}
	return row0 + r;
}
__device__ __forceinline__ bool row_valid(const MatvecArgs& a, int row0, int r, int R) {
	if (a.epi == EPI_GLU) return row0 / 2 + (r % (R / 2)) < a.d;
	return row0 + r < a.d;
}

// TYPE: format tag (an xalm_type, or -1 for the shared-memory LUT path).  R rows per warp, KS warps split K,
// NW warps per CTA.  NORM: fused rmsnorm prologue.
template <int TYPE, int R, int KS, int NW, bool NORM>
__global__ void __launch_bounds__(NW * 32) matvec_kernel(const MatvecArgs a) {
	using F = Fmt<TYPE>;
	constexpr int E = F::E;
	constexpr int GROUPS = NW / KS; // row groups per CTA
	__shared__ float s_part[GROUPS][KS][R];
	__shared__ float s_red[NW];
	__shared__ float s_lut[TYPE == -1 ? 256 : 1];

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int group = warp / KS, ks = warp % KS;
	const int vrows = a.epi == EPI_GLU ? 2 * a.d : a.d; // virtual rows
	const int row0 = (blockIdx.x * GROUPS + group) * R;
	const bool active = row0 < vrows;

	pdl_launch_dependents();
	int tl = -1;
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		tl = tl_begin(200 + a.epi);
		if (a.progress) *reinterpret_cast<volatile unsigned int*>(a.progress) = (unsigned int) a.prog_idx;
	}

	RowPtr rp[R];
#pragma unroll
	for (int r = 0; r < R; r++) {
		const int pr = (active && row_valid(a, row0, r, R)) ? phys_row(a, row0, r, R) : (active ? phys_row(a, row0, 0, R) : 0);
		rp[r] = {a.w.p0 + (size_t) pr * a.w.s0, a.w.p1 + (size_t) pr * a.w.s1, a.w.p2 + (size_t) pr * a.w.s2};
	}
	const int nchunks = a.n / E;
	int c = ks * 32 + lane;

	// ---- first weight fragments are requested before we wait on the previous kernel ----
	typename F::Frag frag[R];
	if (active && c < nchunks) {
#pragma unroll
		for (int r = 0; r < R; r++) frag[r] = F::load(rp[r], c);
	}

	if (TYPE == -1) {
		for (int i = threadIdx.x; i < 256; i += NW * 32) s_lut[i] = decode_byte_type(a.lut_type, (uint8_t) i);
	}

	pdl_wait(); // activations (and the KV ring) written by earlier kernels are visible from here on
	tl_mark(tl, 2);

	if (a.epi == EPI_QKV && blockIdx.x == 0 && a.step->kv_sink > 0) rotate_sinks(a, a.step->kv_sink);

	// ---- rmsnorm prologue: every CTA recomputes the scale (n floats from L1/L2, hidden under the weight fetch) ----
	float scale = 1.0f;
	if (NORM) {
		float ss = 0.f;
		for (int i = threadIdx.x * 4; i < a.n; i += NW * 32 * 4) {
			const float4 v = ld_act4(a.x + i);
			ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
		}
		ss = warp_sum(ss);
		if (lane == 0) s_red[warp] = ss;
		__syncthreads();
		float tot = 0.f;
#pragma unroll
		for (int i = 0; i < NW; i++) tot += s_red[i];
		scale = 1.0f / sqrtf(tot / (float) a.n + a.norm_eps);
	} else if (TYPE == -1) {
		__syncthreads();
	}

	f32x2 acc[R];
#pragma unroll
	for (int r = 0; r < R; r++) acc[r] = pack2(0.f, 0.f);

	if (active) {
		while (c < nchunks) {
			// activations for this chunk
			float xv[E];
#pragma unroll
			for (int i = 0; i < E; i += 4) {
				const float4 v = ld_act4(a.x + (size_t) c * E + i);
				xv[i] = v.x; xv[i + 1] = v.y; xv[i + 2] = v.z; xv[i + 3] = v.w;
			}
			if (NORM) {
				// o[i] = x[i] * scale * weight[i]  (infer.cpp:233-235)
				if (a.norm_type == XALM_F32) {
					const float* g = reinterpret_cast<const float*>(a.norm_w) + (size_t) c * E;
#pragma unroll
					for (int i = 0; i < E; i += 4) {
						const float4 gv = ld_act4(g + i);
						xv[i] = xv[i] * scale * gv.x; xv[i + 1] = xv[i + 1] * scale * gv.y;
						xv[i + 2] = xv[i + 2] * scale * gv.z; xv[i + 3] = xv[i + 3] * scale * gv.w;
					}
				} else {
					const uint16_t* g = reinterpret_cast<const uint16_t*>(a.norm_w) + (size_t) c * E;
#pragma unroll
					for (int i = 0; i < E; i += 4) {
						const uint2 gv = *reinterpret_cast<const uint2*>(g + i);
						xv[i] = xv[i] * scale * __uint_as_float(gv.x << 16); xv[i + 1] = xv[i + 1] * scale * __uint_as_float(gv.x & 0xFFFF0000u);
						xv[i + 2] = xv[i + 2] * scale * __uint_as_float(gv.y << 16); xv[i + 3] = xv[i + 3] * scale * __uint_as_float(gv.y & 0xFFFF0000u);
					}
				}
			}
			// next fragments go out before this chunk's arithmetic
			const int cn = c + KS * 32;
			typename F::Frag nxt[R];
			if (cn < nchunks) {
#pragma unroll
				for (int r = 0; r < R; r++) nxt[r] = F::load(rp[r], cn);
			}
#pragma unroll
			for (int r = 0; r < R; r++) {
				if constexpr (TYPE == -1) F::fma_chunk_lut(frag[r], xv, acc[r], s_lut);
				else F::fma_chunk(frag[r], xv, acc[r]);
			}
			if (cn < nchunks) {
#pragma unroll
				for (int r = 0; r < R; r++) frag[r] = nxt[r];
			}
			c = cn;
		}
	}

	float y[R];
#pragma unroll
	for (int r = 0; r < R; r++) {
		float lo, hi;
		unpack2(acc[r], lo, hi);
		y[r] = warp_sum(lo + hi);
	}
	if (KS == 1) {
		if (active && lane == 0) epilogue<R>(a, row0, y);
	} else {
		if (lane == 0) {
#pragma unroll
			for (int r = 0; r < R; r++) s_part[group][ks][r] = y[r];
		}
		__syncthreads();
		if (active && ks == 0 && lane == 0) {
#pragma unroll
			for (int r = 0; r < R; r++) {
				float t = 0.f;
#pragma unroll
				for (int k = 0; k < KS; k++) t += s_part[group][k][r];
				y[r] = t;
			}
			epilogue<R>(a, row0, y);
		}
	}
	tl_mark(tl, 3);
}

// ---- TQ1_0 (quants.py:664-683): 256 weights in 54 bytes, base-3 packed.  Not a 16-byte-chunk format: one warp
//      walks a row block by block, lane B decodes byte B (5 trits) of qs0, lanes 0-15 byte B of qs1, lanes 0-3 qh. ----
template <int R, int NW, bool NORM>
__global__ void __launch_bounds__(NW * 32) matvec_tq1_kernel(const MatvecArgs a) {
	__shared__ float s_red[NW];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int vrows = a.epi == EPI_GLU ? 2 * a.d : a.d;
	const int row0 = (blockIdx.x * NW + warp) * R;
	const bool active = row0 < vrows;
	pdl_launch_dependents();
	if (blockIdx.x == 0 && threadIdx.x == 0 && a.progress) *reinterpret_cast<volatile unsigned int*>(a.progress) = (unsigned int) a.prog_idx;
	pdl_wait();
	if (a.epi == EPI_QKV && blockIdx.x == 0 && a.step->kv_sink > 0) rotate_sinks(a, a.step->kv_sink);
	float scale = 1.0f;
	if (NORM) {
		float ss = 0.f;
		for (int i = threadIdx.x; i < a.n; i += NW * 32) ss += a.x[i] * a.x[i];
		ss = warp_sum(ss);
		if (lane == 0) s_red[warp] = ss;
		__syncthreads();
		float tot = 0.f;
		for (int i = 0; i < NW; i++) tot += s_red[i];
		scale = 1.0f / sqrtf(tot / (float) a.n + a.norm_eps);
	}
	auto xval = [&](int j) -> float {
		float v = a.x[j];
		if (NORM) v = v * scale * (a.norm_type == XALM_F32 ? reinterpret_cast<const float*>(a.norm_w)[j]
		                                                   : bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(a.norm_w)[j]));
		return v;
	};
	float y[R];
#pragma unroll
	for (int r = 0; r < R; r++) y[r] = 0.f;
	if (active) {
		const int nb = a.n / 256;
		for (int b = 0; b < nb; b++) {
			const int base = b * 256;
#pragma unroll
			for (int r = 0; r < R; r++) {
				const int pr = row_valid(a, row0, r, R) ? phys_row(a, row0, r, R) : phys_row(a, row0, 0, R);
				const uint8_t* blk = a.w.p0 + (size_t) pr * a.w.s0 + (size_t) b * 54;
				float s = 0.f;
				uint8_t q = blk[lane];
#pragma unroll
				for (int k = 0; k < 5; k++) {
					const int trit = ((int) q * 3) >> 8;
					s += (float) (trit - 1) * xval(base + k * 32 + lane);
					q = (uint8_t) (q * 3);
				}
				if (lane < 16) {
					q = blk[32 + lane];
#pragma unroll
					for (int k = 0; k < 5; k++) {
						const int trit = ((int) q * 3) >> 8;
						s += (float) (trit - 1) * xval(base + 160 + k * 16 + lane);
						q = (uint8_t) (q * 3);
					}
				}
				if (lane < 4) {
					q = blk[48 + lane];
#pragma unroll
					for (int k = 0; k < 4; k++) {
						const int trit = ((int) q * 3) >> 8;
						s += (float) (trit - 1) * xval(base + 240 + k * 4 + lane);
						q = (uint8_t) (q * 3);
					}
				}
				y[r] += ld_f16_unaligned(blk + 52) * s;
			}
		}
	}
#pragma unroll
	for (int r = 0; r < R; r++) y[r] = warp_sum(y[r]);
	if (active && lane == 0) epilogue<R>(a, row0, y);
}

} // namespace xalm
