// decode_mega.cu — one persistent kernel per decode token (Model::forward, infer.cpp:604-638, batch 1).
//
// Why.  With one kernel per fused op (5 per layer, 162 per Mistral-7B token) the round-1 timeline
// (profiles/r1_timeline_multikernel.md) shows ~36 us of HBM streaming per layer at roofline against 68 us measured: every
// kernel boundary drains the weight stream (ramp-up, tail, dependency flush, activation staging), and the 8-bit consumer
// loop itself was issue/latency-bound.  Here a token is a list of PHASES run by the same resident CTAs, one per SM:
//
//   [norm+QKV+rope+KV] -> [attention] -> [Wo+residual] -> [norm+W1|W3+GLU] -> [W2+residual] -> ... -> [norm+classifier]
//
//   * one PRODUCER warp per CTA walks the token's whole tile list and keeps a shared-memory ring (up to ~170 KB per SM,
//     ~25 MB chip-wide) full with cp.async.bulk copies signalled through mbarriers.  Weights are immutable, so it never
//     waits for anybody: while the consumers wait for activations or run attention, the ring fills with the NEXT phases'
//     weights and HBM keeps streaming.  Small phases (Wo, QKV) are consumed straight from shared memory.
//   * eight CONSUMER warps per CTA run the integer-dot core (idp.cuh): activations are staged once per phase as block
//     floating point (three int8 limbs per element), a 32-weight block costs 24 dp4a, ~1.3 issue slots per weight, so the
//     consumers outrun the stream and catch up after every hand-off.
//   * phases are separated by a grid-wide hand-off among the consumer warps: __threadfence + relaxed add on a counter in
//     global memory by one thread per CTA, acquire-poll by one thread per CTA, then the activations written by other SMs are
//     read with ld.global.cg (L2), never through L1.  (Tried and measured slower: hand-off by polling {value, tag} words —
//     36 k threads polling the same 32 KB turn its L2 lines into a hot spot, 4-10 us per hand-off; and the same with a
//     fence-free counter as a hint — without the fence the producers' stores take microseconds to reach L2 behind the weight
//     stream, so the pollers spin just as long.  profiles/r2_token_kernel.md.)
//   * everything else a consumer needs after its input arrives is already on the SM: phase descriptors are fetched one
//     phase ahead into shared memory, the token's scalars and the RoPE table are loaded once, norm weights are requested
//     before the input is polled (all of these are read once per token, i.e. from DRAM behind the weight stream).
//   * attention: split-K flash-decode as in attention.cuh, one (kv head, split) item per CTA, first K/V batch requested
//     before the hand-off wait; the merge of the splits is DISTRIBUTED — after one more hand-off CTA c merges outputs
//     [32c, 32c+32) of all splits — instead of a fence + ticket + last-CTA chain.
//   * tile t of a phase runs on CTA (t + tile_off) % grid with tile_off advanced by each phase's remainder, so the odd tile
//     moves around and every SM streams the same number of bytes per token.
//
// Mapping of a matvec phase.  A tile is 8 rows (a RoPE pair / a W1,W3 pair stay adjacent), a ring stage is 8 rows x 16 units
// (4096 elements).  Warp (kw, rw): K-slice kw of 4 (32 blocks of 32 elements, one block per lane) x row group rw of 2
// (4 rows each).  Per stage a lane loads its activation block once (7 shared loads) and walks its 4 rows (2-3 shared loads,
// 24 dp4a and ~8 finishing instructions each).  Lane sums are reduced with a 6-shuffle transposed butterfly, K-slices are
// combined through shared memory in a fixed order (bit-reproducible run to run), a rotating warp runs the epilogue.
#define XALM_SECONDARY_TU
#include "decode_mega.h"

#include <math_constants.h>

#include "idp.cuh"

namespace xalm {

constexpr unsigned int DM_SPIN_LIMIT = 1u << 21; // polls (~0.5-1 us each) before a wait for another SM's data gives up

__device__ __forceinline__ void dm_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Every wait in this kernel is bounded: a wait that gives up raises the CTA's abort flag (later waits return at once), the
// device-wide abort word gbar[1] (other CTAs stop waiting for this one) and the pinned host word the next synchronising call
// turns into an error — a protocol bug or a lost CTA must not hang the GPU.
__device__ __forceinline__ bool dm_mbar_wait(uint64_t* bar, uint32_t parity, volatile int* s_abort) {
	unsigned int spins = 0;
	for (;;) {
		uint32_t ok;
		asm volatile(
		    "{\n"
		    ".reg .pred p;\n"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		    "selp.u32 %0, 1, 0, p;\n"
		    "}\n"
		    : "=r"(ok)
		    : "r"(smem_u32(bar)), "r"(parity)
		    : "memory");
		if (ok) return true;
		if ((++spins & 1023u) == 0) {
			if (*s_abort) return false;
			if (spins > (1u << 20)) { *s_abort = 1; return false; }
		}
	}
}

// ---- tagged words ----------------------------------------------------------------------------------------------------------
struct DmAbort {
	volatile int* s_abort;  // this CTA
	unsigned int* g_abort;  // device-wide
	unsigned int* err;      // pinned host word
};
__device__ __forceinline__ float4 dm_ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
// called every so often from a polling loop: has somebody given up, or is it time to?
__device__ __forceinline__ bool dm_check_abort(const DmAbort ab, unsigned int spins) {
	unsigned int g;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(ab.g_abort) : "memory");
	if (g || *ab.s_abort || spins > DM_SPIN_LIMIT) {
		*ab.s_abort = 1;
		asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(ab.g_abort), "r"(1u) : "memory");
		if (ab.err) *ab.err = 2u;
		return true;
	}
	return false;
}
// grid-wide hand-off among the consumer warps of all CTAs, in two halves so that work which does not depend on the other CTAs
// (norm weights, the first K/V batch, the next descriptor) travels while the counter is polled.  signal: everything this CTA
// wrote is released; wait: everything the others wrote before their signal is acquired.  The counter only grows inside a launch.
__device__ __forceinline__ void dm_signal(unsigned int* ctr) {
	dm_bar(); // all consumer warps of this CTA have issued their stores
	if (threadIdx.x == 0) {
		__threadfence();
		asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
	}
}
__device__ __forceinline__ void dm_wait(unsigned int* ctr, unsigned int target, const DmAbort ab) {
	if (threadIdx.x == 0) {
		unsigned int v, spins = 0;
		for (;;) {
			asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
			if ((int) (v - target) >= 0) break;
			if ((++spins & 255u) == 0 && dm_check_abort(ab, spins)) break;
		}
	}
	dm_bar();
}

// ---- activation staging: x (global, written by other SMs) [-> rmsnorm] -> block floating point in shared memory -----
// Thread t stages 32-element block t (+256, +512, ...): eight 16-byte loads per block, all in flight together.
// norm weights of the thread's (first) block are requested BEFORE the hand-off wait: they are constants, but read once per
// token they come from DRAM, which behind a saturated weight stream is microseconds away.
struct DmNormW {
	uint4 g[8]; // BF16: g[0..3] hold the 32 weights of the thread's block packed; F32: g[0..7] hold them as floats
};
__device__ __forceinline__ void dm_norm_fetch(const MatvecArgs& a, DmNormW& w) {
	const int i = (int) threadIdx.x * 32;
	if (i >= a.n) return;
	if (a.norm_type == XALM_F32) {
#pragma unroll
		for (int c = 0; c < 8; c++) w.g[c] = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(a.norm_w) + i + 4 * c);
	} else {
#pragma unroll
		for (int c = 0; c < 4; c++) w.g[c] = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.norm_w) + i + 8 * c);
	}
}
// norm weight of element e (0..31) of the block whose weights sit in w
__device__ __forceinline__ float dm_norm_get(const MatvecArgs& a, const DmNormW& w, int e) {
	if (a.norm_type == XALM_F32) {
		const uint4 q = w.g[e >> 2];
		const uint32_t u = (e & 3) == 0 ? q.x : (e & 3) == 1 ? q.y : (e & 3) == 2 ? q.z : q.w;
		return __uint_as_float(u);
	}
	const uint4 q = w.g[e >> 3];
	const int h = (e & 7) >> 1;
	const uint32_t u = h == 0 ? q.x : h == 1 ? q.y : h == 2 ? q.z : q.w;
	return __uint_as_float((e & 1) ? (u & 0xFFFF0000u) : (u << 16)); // BF16 = bits << 16, types.h:322-325
}
__device__ __forceinline__ void dm_load_block(const float* x, float (&v)[32]) {
#pragma unroll
	for (int c = 0; c < 8; c++) {
		const float4 q = dm_ld_cg4(x + 4 * c);
		v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
	}
}
// rmsnorm-fused staging (infer.cpp:224-236).  n <= 8192: the thread keeps its block in registers across the sum-of-squares
// reduction.  Longer rows park the raw values in `stash` (the tail of the activation area) and take the norm weights late.
// Out of line on purpose: its 64 registers of block + weights must not weigh on the allocation of the tile loop.  The hand-off
// wait sits inside so that the norm weights travel while the counter is polled.
__device__ __noinline__ void dm_stage_norm(const MatvecArgs& a, uint8_t* xq_base, float* s_red, float* stash, unsigned int* ctr, unsigned int expected,
                                           bool do_wait, const DmAbort ab, unsigned long long* tls) {
	const int n = a.n, nb = n / 32;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const XqView v = xq_view(xq_base, n);
	DmNormW w;
	dm_norm_fetch(a, w);
	if (do_wait) dm_wait(ctr, expected, ab);
	if (tls) tls[1] = gtime();
	float xv[32];
	float ss = 0.f;
	if (tid < nb) {
		dm_load_block(a.x + tid * 32, xv);
#pragma unroll
		for (int e = 0; e < 32; e++) ss += xv[e] * xv[e];
	}
	for (int blk = tid + DM_CW * 32; blk < nb; blk += DM_CW * 32) { // n > 8192
		float t[32];
		dm_load_block(a.x + blk * 32, t);
#pragma unroll
		for (int e = 0; e < 32; e++) { ss += t[e] * t[e]; stash[blk * 32 + e] = t[e]; }
	}
	if (tls) tls[4] = gtime();
	ss = warp_sum(ss);
	if (lane == 0) s_red[warp] = ss;
	dm_bar();
	if (tls) tls[5] = gtime();
	float tot = 0.f;
#pragma unroll
	for (int i = 0; i < DM_CW; i++) tot += s_red[i];
	const float scale = 1.0f / sqrtf(tot / (float) n + a.norm_eps); // infer.cpp:229-232
	if (tid < nb) {
#pragma unroll
		for (int e = 0; e < 32; e++) xv[e] = xv[e] * scale * dm_norm_get(a, w, e); // infer.cpp:233-235
		xq_store_block(v, tid, xv);
	}
	for (int blk = tid + DM_CW * 32; blk < nb; blk += DM_CW * 32) {
		float t[32];
#pragma unroll
		for (int e = 0; e < 32; e++) {
			const int i = blk * 32 + e;
			const float g = a.norm_type == XALM_F32 ? reinterpret_cast<const float*>(a.norm_w)[i] : bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(a.norm_w)[i]);
			t[e] = stash[i] * scale * g;
		}
		xq_store_block(v, blk, t);
	}
}
__device__ __noinline__ void dm_stage_plain(const MatvecArgs& a, uint8_t* xq_base) {
	const XqView v = xq_view(xq_base, a.n);
	const int nb = a.n / 32;
	for (int blk = threadIdx.x; blk < nb; blk += DM_CW * 32) {
		float t[32];
		dm_load_block(a.x + blk * 32, t);
		xq_store_block(v, blk, t);
	}
}

// ---- attention phase -----------------------------------------------------------------------------------------------------------
// Part A: the 8 consumer warps process (kv head, split) items; same math as attn_decode_kernel (attention.cuh).  KVDIV = 2 serves
// 2 x G query heads per kv head as two "virtual" kv heads of G heads each (G = 8 would need > 200 registers per thread; the
// second pass re-reads the K/V slice from L2).  The first K/V batch is requested BEFORE the hand-off wait; rows written during
// this token (kv_pos; the re-rotated sinks) are fetched again after it.
template <int HD, int G, int KVDIV>
__device__ __noinline__ unsigned int dm_attention(const DmPhase& P, const StepParams& st, float* scratch, int first, int stride, const DmAbort ab,
                                                     unsigned int* ctr, unsigned int expected, bool wait_first, unsigned long long* tls) {
	constexpr int NW = DM_CW;
	constexpr int LPR = HD / 8, RPW = 32 / LPR, TB = 4, NGRP = NW * RPW;
	const AttnArgs& a = P.at;
	float* s_m = scratch;                       // [NGRP][G]
	float* s_l = s_m + NGRP * G;                // [NGRP][G]
	float* s_scale = s_l + NGRP * G;            // [NGRP][G]
	float* s_acc = s_scale + NGRP * G;          // [NW][G][HD]
	float* s_q = s_acc + NW * G * HD;           // [G][HD]: q of this item's heads (read back per row: 64 registers less than holding it)
	const int kv_len = st.kv_len, kv_pos = st.kv_pos, kv_sink = st.kv_sink;
	const int slen = attn_split_len(kv_len, a.n_splits, a.min_split);
	const int n_active = (kv_len + slen - 1) / slen;
	const int n_vkv = a.n_kv_heads * KVDIV;
	const int n_items = n_vkv * n_active;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int sub = lane / LPR, li = lane % LPR;
	const int kv_stride = a.n_kv_heads * HD;
	const float inv_sqrt = 1.0f / sqrtf((float) HD);
	for (int item = first; item < n_items; item += stride) {
		const int kvh = item / n_active, split = item % n_active; // kvh: virtual kv head (G query heads each)
		const int kvp = kvh / KVDIV;                              // physical kv head
		const int t0 = split * slen, t1 = min(kv_len, t0 + slen);
		const __half* kbase = a.k_cache + (size_t) kvp * HD + li * 8;
		const __half* vbase = a.v_cache + (size_t) kvp * HD + li * 8;
		uint4 kq[TB], vq[TB];
		auto fetch = [&](int tb) {
#pragma unroll
			for (int j = 0; j < TB; j++) {
				const int t = tb + j * RPW + sub;
				const int tc = t < t1 ? t : t0;
				kq[j] = __ldcg(reinterpret_cast<const uint4*>(kbase + (size_t) tc * kv_stride));
				vq[j] = __ldcg(reinterpret_cast<const uint4*>(vbase + (size_t) tc * kv_stride));
			}
		};
		int tb = t0 + warp * RPW * TB;
		bool have = false;
		if (tb < t1) { fetch(tb); have = true; } // in flight while the QKV phase of the other CTAs is awaited
		if (wait_first) { dm_wait(ctr, expected, ab); wait_first = false; }
		if (tls) tls[4] = gtime();
		for (int i = threadIdx.x * 4; i < G * HD; i += NW * 32 * 4) // q of the item's G heads, parked in shared memory
			*reinterpret_cast<float4*>(s_q + i) = dm_ld_cg4(a.q + (size_t) kvh * G * HD + i);
		dm_bar();
		float m[G], l[G], acc[G][8];
#pragma unroll
		for (int g = 0; g < G; g++) {
			m[g] = -CUDART_INF_F; l[g] = 0.f;
#pragma unroll
			for (int i = 0; i < 8; i++) acc[g][i] = 0.f;
		}
		for (; tb < t1; tb += NW * RPW * TB) {
			if (!have) fetch(tb);
			have = false;
			bool ok[TB];
#pragma unroll
			for (int j = 0; j < TB; j++) {
				const int t = tb + j * RPW + sub;
				ok[j] = t < t1;
				if (ok[j] && (t == kv_pos || t < kv_sink)) { // written during this token: the early request may have seen the old row
					kq[j] = __ldcg(reinterpret_cast<const uint4*>(kbase + (size_t) t * kv_stride));
					vq[j] = __ldcg(reinterpret_cast<const uint4*>(vbase + (size_t) t * kv_stride));
				}
			}
			float s[TB][G];
#pragma unroll
			for (int j = 0; j < TB; j++) {
				const __half2* kh = reinterpret_cast<const __half2*>(&kq[j]);
				float kf[8];
#pragma unroll
				for (int i = 0; i < 4; i++) {
					const float2 f = __half22float2(kh[i]);
					kf[2 * i] = f.x; kf[2 * i + 1] = f.y;
				}
#pragma unroll
				for (int g = 0; g < G; g++) {
					const float4 qa = *reinterpret_cast<const float4*>(s_q + g * HD + li * 8), qb = *reinterpret_cast<const float4*>(s_q + g * HD + li * 8 + 4);
					float p = qa.x * kf[0] + qa.y * kf[1] + qa.z * kf[2] + qa.w * kf[3] + qb.x * kf[4] + qb.y * kf[5] + qb.z * kf[6] + qb.w * kf[7];
#pragma unroll
					for (int o = LPR / 2; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
					s[j][g] = ok[j] ? p * inv_sqrt : -CUDART_INF_F;
				}
			}
#pragma unroll
			for (int g = 0; g < G; g++) {
				float mn = m[g];
#pragma unroll
				for (int j = 0; j < TB; j++) mn = fmaxf(mn, s[j][g]);
				if (mn == -CUDART_INF_F) continue;
				const float corr = expf(m[g] - mn);
				l[g] *= corr;
#pragma unroll
				for (int i = 0; i < 8; i++) acc[g][i] *= corr;
				m[g] = mn;
#pragma unroll
				for (int j = 0; j < TB; j++) {
					const float p = expf(s[j][g] - mn);
					l[g] += p;
					const __half2* vh = reinterpret_cast<const __half2*>(&vq[j]);
#pragma unroll
					for (int i = 0; i < 4; i++) {
						const float2 f = __half22float2(vh[i]);
						acc[g][2 * i] += p * f.x;
						acc[g][2 * i + 1] += p * f.y;
					}
				}
			}
		}
		if (tls) tls[5] = gtime();
		// ---- merge lane groups and warps ----
		const int grp = warp * RPW + sub;
		if (li == 0) {
#pragma unroll
			for (int g = 0; g < G; g++) { s_m[grp * G + g] = m[g]; s_l[grp * G + g] = l[g]; }
		}
		dm_bar();
#pragma unroll
		for (int g = 0; g < G; g++) {
			float M = -CUDART_INF_F;
			for (int i = 0; i < NGRP; i++) M = fmaxf(M, s_m[i * G + g]);
			const float sc = m[g] == -CUDART_INF_F ? 0.f : expf(m[g] - M);
			if (li == 0) s_scale[grp * G + g] = sc;
#pragma unroll
			for (int i = 0; i < 8; i++) {
				float v = acc[g][i] * sc;
#pragma unroll
				for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
				acc[g][i] = v;
			}
			if (sub == 0) {
#pragma unroll
				for (int i = 0; i < 8; i++) s_acc[(warp * G + g) * HD + li * 8 + i] = acc[g][i];
			}
		}
		dm_bar();
		const bool single = n_active == 1; // one split: write the normalised output directly
		float* pacc = a.part_acc + ((size_t) kvh * a.n_splits + split) * G * HD;
		float* pml = a.part_ml + ((size_t) kvh * a.n_splits + split) * G * 2;
		for (int i = threadIdx.x; i < G * HD; i += NW * 32) {
			const int g = i / HD, dpos = i % HD;
			float v = 0.f;
#pragma unroll
			for (int w = 0; w < NW; w++) v += s_acc[(w * G + g) * HD + dpos];
			if (single) {
				float L = 0.f;
				for (int k = 0; k < NGRP; k++) L += s_l[k * G + g] * s_scale[k * G + g];
				a.out[(size_t) kvh * G * HD + i] = v / L;
			} else {
				pacc[i] = v;
			}
		}
		if (!single && threadIdx.x < G) {
			const int g = threadIdx.x;
			float L = 0.f, M = -CUDART_INF_F;
			for (int k = 0; k < NGRP; k++) { L += s_l[k * G + g] * s_scale[k * G + g]; M = fmaxf(M, s_m[k * G + g]); }
			pml[2 * g] = M;
			pml[2 * g + 1] = L;
		}
		dm_bar(); // scratch reuse by the next item / the merge
	}
	// ---- Part B: distributed merge of the splits — CTA c merges outputs [32c, 32c + 32) (one head slice), in split order ----
	if (wait_first) dm_wait(ctr, expected, ab); // no item here: still part of the hand-off (the merge below reads other CTAs' data)
	if (n_active > 1) {
		const int n_out = n_vkv * G * HD;
		if (tls) tls[6] = gtime();
		dm_signal(ctr);
		expected += gridDim.x;
		dm_wait(ctr, expected, ab); // EVERY CTA: one that ran ahead to its end-of-phase signal would be counted in place of a late split
		if (tls) tls[7] = gtime();
		float* s_ms = scratch;                    // [DM_MAX_SPLITS] split maxima of this head
		float* s_num = scratch + DM_MAX_SPLITS;   // [8][32]
		float* s_den = s_num + 8 * 32;            // [8]
		const int ol = threadIdx.x & 31, sg = threadIdx.x >> 5;
		for (int c = (int) blockIdx.x; c * 32 < n_out; c += (int) gridDim.x) {
			const int o = c * 32 + ol;
			const int h = o / HD, d = o % HD, kvh = h / G, g = h % G;
			constexpr int MS = DM_MAX_SPLITS / 8;
			float mv[MS], lv[MS], vv[MS];
#pragma unroll
			for (int i = 0; i < MS; i++) {
				const int sp = sg + 8 * i;
				mv[i] = -CUDART_INF_F; lv[i] = 0.f; vv[i] = 0.f;
				if (sp < n_active) {
					const size_t base = ((size_t) kvh * a.n_splits + sp) * G + g;
					vv[i] = __ldcg(a.part_acc + base * HD + d);
					const float2 ml = __ldcg(reinterpret_cast<const float2*>(a.part_ml + base * 2));
					mv[i] = ml.x; lv[i] = ml.y;
					if (ol == 0) s_ms[sp] = mv[i];
				}
			}
			dm_bar();
			float M = -CUDART_INF_F;
			for (int sp = 0; sp < n_active; sp++) M = fmaxf(M, s_ms[sp]);
			float num = 0.f, den = 0.f;
#pragma unroll
			for (int i = 0; i < MS; i++) {
				if (sg + 8 * i < n_active) {
					const float sc = expf(mv[i] - M);
					num += sc * vv[i];
					den += sc * lv[i];
				}
			}
			s_num[sg * 32 + ol] = num;
			if (ol == 0) s_den[sg] = den;
			dm_bar();
			if (sg == 0) {
				float tn = 0.f, td = 0.f;
#pragma unroll
				for (int k = 0; k < 8; k++) { tn += s_num[k * 32 + ol]; td += s_den[k]; }
				a.out[o] = tn / td;
			}
			dm_bar();
		}
	}
	return expected;
}
template <int G, int KVDIV>
__device__ __forceinline__ unsigned int dm_attention_hd(const DmPhase& P, const StepParams& st, float* scratch, int first, int stride, const DmAbort ab,
                                                        unsigned int* ctr, unsigned int expected, bool wait_first, unsigned long long* tls) {
	if (P.HD == 64) return dm_attention<64, G, KVDIV>(P, st, scratch, first, stride, ab, ctr, expected, wait_first, tls);
	return dm_attention<128, G, KVDIV>(P, st, scratch, first, stride, ab, ctr, expected, wait_first, tls);
}

// ---- the tile loop of a matvec phase, out of line: its register allocation is its own (inlined into the phase loop, ptxas kept
//      the ring cursor and the activation pointers in local memory and reloaded them every stage) ----------------------------
struct DmTileCtx { // lives in shared memory
	uint8_t* xq_base;
	uint8_t* ring;
	float* part;
	uint64_t* full;
	uint64_t* empty;
	int* s_abort;
	const StepParams* step;
	const float2* rope;
	int slot_bytes, NS;
};
template <int TYPE>
__device__ __forceinline__ uint32_t dm_tiles(const DmPhase& P, const DmTileCtx& c, int first, uint32_t cursor, unsigned long long* tlp) {
	using F = IdpFmt<TYPE>;
	constexpr int KW = DM_KW, R = DM_R, RC = DM_RC, U = DM_U, UB = F::UB;
	constexpr int ROW_STAGE = U * UB;
	const MatvecArgs& a = P.a;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int kw = warp % KW, rw = warp / KW;
	const int hA = (lane >> 2) & 1;
	const int G = (int) gridDim.x;
	const int n = a.n, nu = n / 256, epi = a.epi;
	const int kranges = P.kranges, n_tiles = P.n_tiles;
	float* const out = a.out;
	const XqView xv = xq_view(c.xq_base, n);
	uint8_t* const ring = c.ring;
	float* const part = c.part;
	uint64_t* const full = c.full;
	uint64_t* const empty = c.empty;
	int* const s_abort_p = c.s_abort;
	const int NS = c.NS, slot_bytes = c.slot_bytes;
	const StepParams& s_step = *c.step;
	const float2* const s_rope = c.rope;
	int tcount = 0;
	int rslot = (int) (cursor >> 1), rphase = (int) (cursor & 1u);
	for (int tile = first; tile < n_tiles; tile += G, tcount++) {
		const int row0 = tile * RC;
		const bool reducer = warp == (tcount % DM_CW);
		// residual: request the old activation early so the epilogue does not sit on a trip to L2
		float xold = 0.f;
		if (epi == EPI_RESIDUAL && reducer && lane < RC) xold = __ldcg(out + row0 + lane);
		float y[R];
#pragma unroll
		for (int r = 0; r < R; r++) y[r] = 0.f;
		for (int kr = 0; kr < kranges; kr++) {
			const int u0 = kr * U;
			const int nb = 8 * min(U, nu - u0); // blocks per row in this stage
			const int b = kw * 32 + lane;
			const bool got = dm_mbar_wait(&full[rslot], rphase, s_abort_p);
			if (tlp && tcount == 0 && kr == 0) tlp[6] = gtime();
			if (got && b < nb) {
				XqBlock xb;
				xq_load(xv, u0 * 8 + b, hA, xb);
				const uint8_t* unit = ring + (size_t) rslot * slot_bytes + (size_t) (rw * R) * ROW_STAGE + (size_t) (b >> 3) * UB;
#pragma unroll
				for (int r = 0; r < R; r++) F::block(unit + (size_t) r * ROW_STAGE, b & 7, hA, xb, y[r]);
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty[rslot]);
			if (++rslot == NS) { rslot = 0; rphase ^= 1; }
		}
		// ---- lanes -> one sum per row (transposed butterfly: 6 shuffles for 4 rows), K-slices -> shared memory (fixed order) ----
		{
			const bool b4 = lane & 16, b3 = lane & 8;
			float k0 = b4 ? y[2] : y[0], k1 = b4 ? y[3] : y[1];
			const float s0 = b4 ? y[0] : y[2], s1 = b4 ? y[1] : y[3];
			k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
			k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
			float k = b3 ? k1 : k0;
			const float s = b3 ? k0 : k1;
			k += __shfl_xor_sync(0xffffffffu, s, 8);
			k += __shfl_xor_sync(0xffffffffu, k, 4);
			k += __shfl_xor_sync(0xffffffffu, k, 2);
			k += __shfl_xor_sync(0xffffffffu, k, 1);
			if ((lane & 7) == 0) part[(tcount & 1) * (KW * RC) + kw * RC + rw * R + (b4 ? 2 : 0) + (b3 ? 1 : 0)] = k;
		}
		dm_bar();
		if (reducer) { // rotating reducer warp: lane i owns row i of the tile
			const float* pt = part + (tcount & 1) * (KW * RC);
			float yv = 0.f;
			if (lane < RC) {
#pragma unroll
				for (int k = 0; k < KW; k++) yv += pt[k * RC + lane];
			}
			const float ynext = __shfl_down_sync(0xffffffffu, yv, 1);
			const int row = row0 + lane;
			if (lane < RC) {
				if (epi == EPI_RESIDUAL) {
					out[row] = xold + yv; // infer.cpp:450-452, :492-494
				} else if (epi == EPI_GLU) {
					if ((lane & 1) == 0) { // (W1[o], W3[o]) sit in adjacent rows of a GLU tile
						const float g = a.act == XALM_SILU ? act_silu(yv) : act_gelu(yv);
						out[row >> 1] = g * ynext; // infer.cpp:470-488
					}
				} else if (epi == EPI_STORE) {
					if (row < a.d) out[row] = yv; // logits (infer.cpp:637)
				} else if ((lane & 1) == 0) { // EPI_QKV: clip -> RoPE -> q, or fp16 K/V into the cache ring (infer.cpp:388-414)
					float v0 = clipf(yv, a.qkv_clip), v1 = clipf(ynext, a.qkv_clip);
					if (row < a.q_dim) {
						const float2 cs = s_rope[(row % a.head_dim) >> 1];
						*reinterpret_cast<float2*>(out + row) = make_float2(v0 * cs.x - v1 * cs.y, v0 * cs.y + v1 * cs.x);
					} else if (row < a.q_dim + a.kv_dim) {
						const int i = row - a.q_dim;
						const float2 cs = s_rope[(i % a.head_dim) >> 1];
						const float r0 = v0 * cs.x - v1 * cs.y, r1 = v0 * cs.y + v1 * cs.x;
						v0 = r0; v1 = r1;
						*reinterpret_cast<__half2*>(a.k_cache + (size_t) s_step.kv_pos * a.kv_dim + i) = __floats2half2_rn(v0, v1);
					} else {
						const int i = row - a.q_dim - a.kv_dim;
						*reinterpret_cast<__half2*>(a.v_cache + (size_t) s_step.kv_pos * a.kv_dim + i) = __floats2half2_rn(v0, v1);
					}
				}
			}
		}
	}
	return ((uint32_t) rslot << 1) | (uint32_t) rphase;
}

template <int TYPE>
__global__ void __launch_bounds__(DM_THREADS, 1) decode_token_kernel(const DmArgs mk) {
	using F = IdpFmt<TYPE>;
	constexpr int KW = DM_KW, R = DM_R, RC = DM_RC, U = DM_U, UB = F::UB;
	constexpr int ROW_STAGE = U * UB; // bytes of one row inside a ring slot
	constexpr int PH_WORDS = (int) (sizeof(DmPhase) / 4);
	static_assert(sizeof(DmPhase) % 4 == 0 && PH_WORDS <= DM_CW * 32, "one descriptor word per consumer thread");
	const int NS = mk.NS;
	const int G = (int) gridDim.x;

	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t* xq_base = smem;
	uint8_t* ring = smem + mk.xq_cap;
	float* part = reinterpret_cast<float*>(ring + (size_t) NS * mk.slot_bytes); // [2][KW][RC]
	float* s_red = part + 2 * KW * RC;                                           // [16]
	uint64_t* full = reinterpret_cast<uint64_t*>(s_red + 16);
	uint64_t* empty = full + NS;
	__shared__ __align__(16) DmPhase s_ph[2]; // this phase's descriptor and the next one's (fetched a phase ahead)
	__shared__ StepParams s_step;
	__shared__ float s_freq[DM_MAX_HD / 2];
	__shared__ float2 s_rope[DM_MAX_HD / 2]; // {cos, sin}(pos * freq[j]): the QKV epilogue's rotation (infer.cpp:305-322) without a libm call per pair
	__shared__ int s_abort;
	__shared__ DmTileCtx s_ctx;
	__shared__ int s_quiet; // consumers are between phases (hand-off, staging, attention): the producer holds its copies back (mk.quiet)

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) {
		s_abort = 0;
		s_quiet = 0;
		s_ctx = {xq_base, ring, part, full, empty, &s_abort, &s_step, s_rope, mk.slot_bytes, NS};
		for (int s = 0; s < NS; s++) {
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], DM_CW);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < PH_WORDS) reinterpret_cast<uint32_t*>(&s_ph[0])[threadIdx.x] = reinterpret_cast<const uint32_t*>(&mk.phases[0])[threadIdx.x];
	if (threadIdx.x < (int) (sizeof(StepParams) / 4)) reinterpret_cast<uint32_t*>(&s_step)[threadIdx.x] = reinterpret_cast<const uint32_t*>(mk.step)[threadIdx.x];
	if (threadIdx.x < mk.head_dim / 2) {
		const float f = mk.rope_freq[threadIdx.x];
		s_freq[threadIdx.x] = f;
		float fci, fcr;
		sincosf((float) mk.step->pos * f, &fci, &fcr); // the same expression as rope_pair (matvec.cuh)
		s_rope[threadIdx.x] = make_float2(fcr, fci);
	}
	__syncthreads();

	if (warp == DM_CW) {
		// ===================== producer: the whole token's weight stream, never blocked by anybody =====================
		if (lane == 0) {
			int slot = 0, phase = 0;
			for (int ph = 0; ph < mk.n_phases; ph++) {
				const DmPhase& P = mk.phases[ph];
				for (int k = ph + 2; k <= ph + 3 && k < mk.n_phases; k++) { // descriptors two and three phases ahead -> L2
					const char* d = reinterpret_cast<const char*>(&mk.phases[k]);
					for (int o = 0; o < (int) sizeof(DmPhase); o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(d + o));
				}
				if (P.kind != DM_MATVEC) continue;
				const int nu = P.a.n / 256, n_tiles = P.n_tiles, kranges = P.kranges;
				const int epi = P.a.epi, glu_off = P.a.glu_off;
				const uint8_t* w0 = P.a.w.p0;
				const size_t ws = P.a.w.s0;
				const int first = ((int) blockIdx.x + G - P.tile_off % G) % G;
				for (int tile = first; tile < n_tiles; tile += G) {
					const int row0 = tile * RC;
					for (int kr = 0; kr < kranges; kr++) {
						if (!dm_mbar_wait(&empty[slot], phase ^ 1, &s_abort)) return;
						if (mk.quiet) { // loads from L2 take ~2 us behind a saturated copy stream: leave the memory system to the consumers' hand-off
							while (*reinterpret_cast<volatile int*>(&s_quiet)) {
								if (*reinterpret_cast<volatile int*>(&s_abort)) return;
								__nanosleep(64);
							}
						}
						const int u0 = kr * U;
						const int un = min(U, nu - u0);
						const uint32_t bytes = (uint32_t) un * UB;
						mbar_expect_tx(&full[slot], bytes * RC);
						uint8_t* dst = ring + (size_t) slot * mk.slot_bytes;
#pragma unroll
						for (int r = 0; r < RC; r++) {
							int pr = row0 + r;
							if (epi == EPI_GLU) { // interleave: even r -> W1[o], odd r -> W3[o]
								const int o = (row0 >> 1) + (r >> 1);
								pr = (r & 1) ? glu_off + o : o;
							}
							bulk_g2s(dst + (size_t) r * ROW_STAGE, w0 + (size_t) pr * ws + (size_t) u0 * UB, bytes, &full[slot]);
						}
						if (++slot == NS) { slot = 0; phase ^= 1; }
					}
				}
			}
		}
		return;
	}

	// ===================== consumers =====================
	uint32_t cursor = 0; // ring cursor of the consumers: (slot << 1) | parity
	const DmAbort ab = {&s_abort, mk.gbar + 1, mk.err};
	unsigned int expected = 0; // arrivals on the hand-off counter so far (same sequence in every CTA)
	unsigned long long* tl = (mk.tl && blockIdx.x == 0 && threadIdx.x == 0) ? mk.tl : nullptr;

	for (int ph = 0; ph < mk.n_phases; ph++) {
		const DmPhase& P = s_ph[ph & 1];
		if (tl) tl[8 * ph] = gtime();
		if (mk.tl && threadIdx.x == 0) mk.tl[(size_t) 8 * mk.tl_phases + (size_t) ph * G + blockIdx.x] = gtime();
		// the next phase's descriptor travels while this phase runs (read once per token: DRAM, or L2 thanks to the producer)
		uint32_t next_word = 0;
		const bool has_next = ph + 1 < mk.n_phases && threadIdx.x < PH_WORDS;
		if (has_next) next_word = reinterpret_cast<const uint32_t*>(&mk.phases[ph + 1])[threadIdx.x];
		const int first = ((int) blockIdx.x + G - P.tile_off % G) % G;
		if (P.kind == DM_ATTN) {
			if (tl) tl[8 * ph + 1] = tl[8 * ph];
			float* scratch = reinterpret_cast<float*>(xq_base);
			switch (P.G) {
				case 1: expected = dm_attention_hd<1, 1>(P, s_step, scratch, first, G, ab, mk.gbar, expected, ph > 0, tl ? tl + 8 * ph : nullptr); break;
				case 2: expected = dm_attention_hd<2, 1>(P, s_step, scratch, first, G, ab, mk.gbar, expected, ph > 0, tl ? tl + 8 * ph : nullptr); break;
				case 4: expected = dm_attention_hd<4, 1>(P, s_step, scratch, first, G, ab, mk.gbar, expected, ph > 0, tl ? tl + 8 * ph : nullptr); break;
				case 8: expected = dm_attention_hd<4, 2>(P, s_step, scratch, first, G, ab, mk.gbar, expected, ph > 0, tl ? tl + 8 * ph : nullptr); break;
			}
			if (tl) { tl[8 * ph + 2] = tl[8 * ph + 1]; tl[8 * ph + 3] = gtime(); }
		} else {
			const MatvecArgs& a = P.a;
			const int n = a.n, nu = n / 256, epi = a.epi;
			const int kranges = P.kranges, n_tiles = P.n_tiles;
			float* const out = a.out;
			const bool norm = a.norm_w != nullptr;
			if (!norm || first >= n_tiles) { // (the norm-fused staging waits inside, after requesting its weights)
				if (ph > 0) dm_wait(mk.gbar, expected, ab);
				if (tl) tl[8 * ph + 1] = gtime();
			}
			if (epi == EPI_QKV && blockIdx.x == 0 && s_step.kv_sink > 0) { // attention sinks move on by one position (infer.cpp:416-431)
				const int pairs = a.kv_dim / 2;
				for (int i = threadIdx.x; i < s_step.kv_sink * pairs; i += DM_CW * 32) {
					const int r = i / pairs, p = i % pairs;
					__half2* kp = reinterpret_cast<__half2*>(a.k_cache + (size_t) r * a.kv_dim) + p;
					float2 vv = __half22float2(__ldcg(kp));
					rope_pair(vv.x, vv.y, (2 * p) % a.head_dim, 1, s_freq);
					*kp = __floats2half2_rn(vv.x, vv.y);
				}
			}
			if (first < n_tiles) {
				if (norm) dm_stage_norm(a, xq_base, s_red, reinterpret_cast<float*>(xq_base + mk.xq_cap) - n, mk.gbar, expected, ph > 0, ab, tl ? tl + 8 * ph : nullptr); // host: xq_cap >= xq_bytes(n) + 4 n
				else dm_stage_plain(a, xq_base);
				dm_bar();
				if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(&s_quiet) = 0;
				if (tl) tl[8 * ph + 2] = gtime();

				cursor = dm_tiles<TYPE>(P, s_ctx, first, cursor, tl ? tl + 8 * ph : nullptr);
			}
			if (tl) tl[8 * ph + 3] = gtime();
		}
		if (has_next) reinterpret_cast<uint32_t*>(&s_ph[(ph + 1) & 1])[threadIdx.x] = next_word;
		if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(&s_quiet) = 1;
		dm_signal(mk.gbar); // (its barrier: every warp is done with this phase's staged activations, partial sums and descriptor)
		expected += (unsigned int) G;
	}
	if (threadIdx.x == 0 && s_abort && mk.err) *mk.err = 3u;
}

bool dm_supported_type(int type) { return idp_supported(type); }
size_t dm_xq_bytes(int n) { return xq_bytes(n); }

size_t dm_attn_scratch_bytes(int HD, int G) {
	if (G == 8) G = 4; // served as two passes of four heads
	const int LPR = HD / 8, RPW = 32 / LPR, NGRP = DM_CW * RPW;
	return ((size_t) 3 * NGRP * G + (size_t) DM_CW * G * HD + (size_t) G * HD) * sizeof(float);
}

size_t dm_fixed_smem(size_t xq_cap, int NS) {
	return xq_cap + 2 * DM_KW * DM_RC * sizeof(float) + 16 * sizeof(float) + 2 * (size_t) NS * sizeof(uint64_t) + 128;
}

template <int TYPE>
static cudaError_t dm_launch_typed(const DmArgs& args, int grid, size_t smem, cudaStream_t s, bool coop) {
	auto kern = decode_token_kernel<TYPE>;
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (e != cudaSuccess) return e;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(DM_THREADS);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeCooperative; // every CTA must be resident: the hand-offs spin on one another
	attr[0].val.cooperative = coop ? 1 : 0;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, args);
}

cudaError_t dm_launch(int type, const DmArgs& args, int grid, size_t smem, cudaStream_t s, bool coop) {
	switch (type) {
		case XALM_Q8_0: return dm_launch_typed<XALM_Q8_0>(args, grid, smem, s, coop);
		case XALM_Q8: return dm_launch_typed<XALM_Q8>(args, grid, smem, s, coop);
		case XALM_Q4_0: return dm_launch_typed<XALM_Q4_0>(args, grid, smem, s, coop);
		case XALM_Q4_1: return dm_launch_typed<XALM_Q4_1>(args, grid, smem, s, coop);
		case XALM_Q5_0: return dm_launch_typed<XALM_Q5_0>(args, grid, smem, s, coop);
		case XALM_Q5_1: return dm_launch_typed<XALM_Q5_1>(args, grid, smem, s, coop);
		
		
		
		
		
	}
	return cudaErrorInvalidValue;
}

} // namespace xalm
