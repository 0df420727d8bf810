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
//   * there is NO grid barrier.  Every vector that crosses SMs (x, q, this token's K/V row, attention partials, xb2, hb) is
//     "tagged": one 64-bit word per element, {value, tag of the producing phase}, written with one 8-byte store by the
//     epilogue that computes it and polled by the consumer that stages it (ld.relaxed.gpu, L2) — the hand-off IS the data:
//     no fence, no flag, no counter, one trip through L2 (measured: a trip costs ~1 us under a saturated weight stream, and
//     a fence + counter + reload hand-off cost three of them, 5 times per layer).  A phase may overwrite a tagged vector in
//     place because every reader of the old value has tiles the writer's own input depends on; a CTA without work in a
//     phase does not read at all, and a reader accepts a NEWER tag, so nobody can wait for a tag that has been overwritten.
//   * everything else a consumer needs after its input arrives is already on the SM: phase descriptors are fetched one
//     phase ahead into shared memory, the token's scalars and the RoPE table are loaded once, norm weights are requested
//     before the input is polled (all of these are read once per token, i.e. from DRAM behind the weight stream).
//   * attention: split-K flash-decode as in attention.cuh, one (kv head, split) item per CTA, first K/V batch requested
//     before q is polled; per-split partials are tagged, and the merge is DISTRIBUTED — CTA c merges outputs [32c, 32c+32)
//     of all splits — instead of a fence + ticket + last-CTA chain.
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
__device__ __forceinline__ dm_tagged dm_pack(uint32_t bits, uint32_t tag) { return (dm_tagged) bits | ((dm_tagged) tag << 32); }
__device__ __forceinline__ dm_tagged dm_packf(float v, uint32_t tag) { return dm_pack(__float_as_uint(v), tag); }
__device__ __forceinline__ bool dm_fresh(dm_tagged w, uint32_t tag) { return (int) ((uint32_t) (w >> 32) - tag) >= 0; }
__device__ __forceinline__ void dm_st(dm_tagged* p, dm_tagged w) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory"); }
__device__ __forceinline__ void dm_st2(dm_tagged* p, dm_tagged w0, dm_tagged w1) {
	asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ dm_tagged dm_ld(const dm_tagged* p) {
	dm_tagged w;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
	return w;
}
__device__ __forceinline__ void dm_ld2(const dm_tagged* p, dm_tagged& a, dm_tagged& b) {
	asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// called every so often from a polling loop: has somebody given up, or is it time to?
__device__ __forceinline__ bool dm_check_abort(const DmAbort& ab, unsigned int spins) {
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
// four consecutive tagged words (32-byte aligned) -> their values, once all four carry `tag` (or a newer one)
struct DmQuad {
	dm_tagged w[4];
};
__device__ __forceinline__ void dm_quad_issue(const dm_tagged* p, DmQuad& q) {
	dm_ld2(p, q.w[0], q.w[1]);
	dm_ld2(p + 2, q.w[2], q.w[3]);
}
__device__ __forceinline__ void dm_quad_wait(const dm_tagged* p, uint32_t tag, DmQuad& q, const DmAbort& ab) {
	unsigned int spins = 0;
	while (!(dm_fresh(q.w[0], tag) && dm_fresh(q.w[1], tag) && dm_fresh(q.w[2], tag) && dm_fresh(q.w[3], tag))) {
		if ((++spins & 63u) == 0 && dm_check_abort(ab, spins)) break;
		dm_quad_issue(p, q);
	}
}
__device__ __forceinline__ float4 dm_quad_f4(const DmQuad& q) {
	return make_float4(__uint_as_float((uint32_t) q.w[0]), __uint_as_float((uint32_t) q.w[1]), __uint_as_float((uint32_t) q.w[2]),
	                   __uint_as_float((uint32_t) q.w[3]));
}
__device__ __forceinline__ dm_tagged dm_poll1(const dm_tagged* p, uint32_t tag, const DmAbort& ab) {
	dm_tagged w = dm_ld(p);
	unsigned int spins = 0;
	while (!dm_fresh(w, tag)) {
		if ((++spins & 63u) == 0 && dm_check_abort(ab, spins)) break;
		w = dm_ld(p);
	}
	return w;
}

// ---- activation staging: tagged x (global, written by other SMs) [-> rmsnorm] -> block floating point in shared memory -----
// norm weights of this thread's chunks (elements tid*4 + c*1024 ..+3), requested BEFORE the input is polled: they are
// constants, but read once per token they come from DRAM, which behind a saturated weight stream is microseconds away.
constexpr int DM_GC = 8; // chunks per thread covered by the early fetch (n <= 8192); longer rows fetch the rest late
struct DmNormW {
	uint2 g[DM_GC]; // BF16 weights, packed (the usual case: convert.py keeps 1-D tensors bf16); F32 weights are fetched late
};
__device__ __forceinline__ uint4 dm_norm_unpack(uint2 packed) { // BF16 = bits << 16, types.h:322-325
	return make_uint4(packed.x << 16, packed.x & 0xFFFF0000u, packed.y << 16, packed.y & 0xFFFF0000u);
}
__device__ __forceinline__ uint4 dm_norm_w4(const MatvecArgs& a, int i) {
	if (a.norm_type == XALM_F32) return *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(a.norm_w) + i);
	return dm_norm_unpack(*reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(a.norm_w) + i));
}
__device__ __forceinline__ void dm_norm_fetch(const MatvecArgs& a, DmNormW& w) {
	if (a.norm_type == XALM_F32) { // pull the lines towards L2; the values are read after the reduction
#pragma unroll
		for (int c = 0; c < DM_GC; c++) {
			const int i = (int) threadIdx.x * 4 + c * (DM_CW * 32 * 4);
			if (i < a.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const float*>(a.norm_w) + i));
		}
		return;
	}
#pragma unroll
	for (int c = 0; c < DM_GC; c++) {
		const int i = (int) threadIdx.x * 4 + c * (DM_CW * 32 * 4);
		if (i < a.n) w.g[c] = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(a.norm_w) + i);
	}
}
// rmsnorm-fused staging (infer.cpp:224-236): pass 1 polls x (all of a thread's chunks in flight together), accumulates the
// sum of squares and parks the raw values in `stash` (the tail of the activation area, free while n = dim is being staged);
// pass 2 scales, multiplies by the norm weight and quantises.
__device__ __forceinline__ void dm_stage_norm(const MatvecArgs& a, const dm_tagged* in_t, uint32_t tag, const XqView& v, float* s_red,
                                              float* stash, const DmNormW& w, const DmAbort& ab) {
	const int n = a.n;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int B = 4;
	float ss = 0.f;
	for (int i0 = tid * 4; i0 < n; i0 += B * DM_CW * 32 * 4) {
		DmQuad q[B];
#pragma unroll
		for (int c = 0; c < B; c++) {
			const int i = i0 + c * (DM_CW * 32 * 4);
			if (i < n) dm_quad_issue(in_t + i, q[c]);
		}
#pragma unroll
		for (int c = 0; c < B; c++) {
			const int i = i0 + c * (DM_CW * 32 * 4);
			if (i < n) {
				dm_quad_wait(in_t + i, tag, q[c], ab);
				const float4 x = dm_quad_f4(q[c]);
				ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
				*reinterpret_cast<float4*>(stash + i) = x;
			}
		}
	}
	ss = warp_sum(ss);
	if (lane == 0) s_red[warp] = ss;
	dm_bar();
	float tot = 0.f;
#pragma unroll
	for (int i = 0; i < DM_CW; i++) tot += s_red[i];
	const float scale = 1.0f / sqrtf(tot / (float) n + a.norm_eps); // infer.cpp:229-232
#pragma unroll
	for (int c = 0; c < DM_GC; c++) { // same elements this thread parked above
		const int i = tid * 4 + c * (DM_CW * 32 * 4);
		if (i < n) {
			const float4 x = *reinterpret_cast<const float4*>(stash + i);
			const uint4 g = a.norm_type == XALM_F32 ? dm_norm_w4(a, i) : dm_norm_unpack(w.g[c]);
			float4 o;
			o.x = x.x * scale * __uint_as_float(g.x); // infer.cpp:233-235
			o.y = x.y * scale * __uint_as_float(g.y);
			o.z = x.z * scale * __uint_as_float(g.z);
			o.w = x.w * scale * __uint_as_float(g.w);
			xq_store4(v, i, o, lane);
		}
	}
	for (int i = tid * 4 + DM_GC * (DM_CW * 32 * 4); i < n; i += DM_CW * 32 * 4) { // rows longer than 8192
		const float4 x = *reinterpret_cast<const float4*>(stash + i);
		const uint4 g = dm_norm_w4(a, i);
		float4 o;
		o.x = x.x * scale * __uint_as_float(g.x);
		o.y = x.y * scale * __uint_as_float(g.y);
		o.z = x.z * scale * __uint_as_float(g.z);
		o.w = x.w * scale * __uint_as_float(g.w);
		xq_store4(v, i, o, lane);
	}
}
__device__ __forceinline__ void dm_stage_plain(const MatvecArgs& a, const dm_tagged* in_t, uint32_t tag, const XqView& v, const DmAbort& ab) {
	const int n = a.n;
	const int tid = threadIdx.x, lane = tid & 31;
	constexpr int B = 8; // chunks in flight per thread: a 7B-class W2 input (14336) takes two trips to L2
	for (int i0 = tid * 4; i0 < n; i0 += B * DM_CW * 32 * 4) {
		DmQuad q[B];
#pragma unroll
		for (int c = 0; c < B; c++) {
			const int i = i0 + c * (DM_CW * 32 * 4);
			if (i < n) dm_quad_issue(in_t + i, q[c]);
		}
#pragma unroll
		for (int c = 0; c < B; c++) {
			const int i = i0 + c * (DM_CW * 32 * 4);
			if (i < n) {
				dm_quad_wait(in_t + i, tag, q[c], ab);
				xq_store4(v, i, dm_quad_f4(q[c]), lane);
			}
		}
	}
}

// ---- attention phase -----------------------------------------------------------------------------------------------------------
// Part A: the 8 consumer warps process (kv head, split) items; same math as attn_decode_kernel (attention.cuh).  KVDIV = 2 serves
// 2 x G query heads per kv head as two "virtual" kv heads of G heads each (G = 8 would need > 200 registers per thread; the
// second pass re-reads the K/V slice from L2).  Rows written during this token (kv_pos; the re-rotated sinks) come from the
// tagged side buffers, everything else from the cache.
template <int HD, int G, int KVDIV>
__device__ __noinline__ void dm_attention(const DmPhase& P, const StepParams& st, float* scratch, int first, int stride, uint32_t tag_in,
                                             uint32_t tag_out, const DmAbort& ab) {
	constexpr int NW = DM_CW;
	constexpr int LPR = HD / 8, RPW = 32 / LPR, TB = 4, NGRP = NW * RPW;
	const AttnArgs& a = P.at;
	float* s_m = scratch;                       // [NGRP][G]
	float* s_l = s_m + NGRP * G;                // [NGRP][G]
	float* s_scale = s_l + NGRP * G;            // [NGRP][G]
	float* s_acc = s_scale + NGRP * G;          // [NW][G][HD]
	const int kv_len = st.kv_len, kv_pos = st.kv_pos, kv_sink = st.kv_sink;
	const int slen = attn_split_len(kv_len, a.n_splits, a.min_split);
	const int n_active = (kv_len + slen - 1) / slen;
	const int n_vkv = a.n_kv_heads * KVDIV;
	const int n_items = n_vkv * n_active;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int sub = lane / LPR, li = lane % LPR;
	const int kv_stride = a.n_kv_heads * HD;
	const int kvd2 = kv_stride / 2;
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
		if (tb < t1) { fetch(tb); have = true; } // in flight while q is polled
		float qf[G][8];
#pragma unroll
		for (int g = 0; g < G; g++) {
			const dm_tagged* qp = P.in_t + (size_t) (kvh * G + g) * HD + li * 8;
			DmQuad q0, q1;
			dm_quad_issue(qp, q0);
			dm_quad_issue(qp + 4, q1);
			dm_quad_wait(qp, tag_in, q0, ab);
			dm_quad_wait(qp + 4, tag_in, q1, ab);
			const float4 u = dm_quad_f4(q0), v = dm_quad_f4(q1);
			qf[g][0] = u.x; qf[g][1] = u.y; qf[g][2] = u.z; qf[g][3] = u.w;
			qf[g][4] = v.x; qf[g][5] = v.y; qf[g][6] = v.z; qf[g][7] = v.w;
		}
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
				if (ok[j] && (t == kv_pos || t < kv_sink)) { // written during this token: take the tagged copy
					const dm_tagged* kp = (t == kv_pos ? P.tkv : P.tsink + (size_t) t * kvd2) + (kvp * HD + li * 8) / 2;
					DmQuad q;
					dm_quad_issue(kp, q);
					dm_quad_wait(kp, tag_in, q, ab);
					kq[j] = make_uint4((uint32_t) q.w[0], (uint32_t) q.w[1], (uint32_t) q.w[2], (uint32_t) q.w[3]);
					if (t == kv_pos) {
						const dm_tagged* vp = P.tkv + kvd2 + (kvp * HD + li * 8) / 2;
						dm_quad_issue(vp, q);
						dm_quad_wait(vp, tag_in, q, ab);
						vq[j] = make_uint4((uint32_t) q.w[0], (uint32_t) q.w[1], (uint32_t) q.w[2], (uint32_t) q.w[3]);
					}
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
					float p = 0.f;
#pragma unroll
					for (int i = 0; i < 8; i++) p += qf[g][i] * kf[i];
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
		dm_tagged* pacc = P.tpart + ((size_t) kvh * a.n_splits + split) * G * HD;
		dm_tagged* pml = P.tml + ((size_t) kvh * a.n_splits + split) * G * 2;
		for (int i = threadIdx.x; i < G * HD; i += NW * 32) {
			const int g = i / HD, dpos = i % HD;
			float v = 0.f;
#pragma unroll
			for (int w = 0; w < NW; w++) v += s_acc[(w * G + g) * HD + dpos];
			if (single) {
				float L = 0.f;
				for (int k = 0; k < NGRP; k++) L += s_l[k * G + g] * s_scale[k * G + g];
				const float o = v / L;
				dm_st(P.out_t + (size_t) kvh * G * HD + i, dm_packf(o, tag_out));
				a.out[(size_t) kvh * G * HD + i] = o;
			} else {
				dm_st(pacc + i, dm_packf(v, tag_out));
			}
		}
		if (!single && threadIdx.x < G) {
			const int g = threadIdx.x;
			float L = 0.f, M = -CUDART_INF_F;
			for (int k = 0; k < NGRP; k++) { L += s_l[k * G + g] * s_scale[k * G + g]; M = fmaxf(M, s_m[k * G + g]); }
			dm_st2(pml + 2 * g, dm_packf(M, tag_out), dm_packf(L, tag_out));
		}
		dm_bar(); // scratch reuse by the next item / the merge
	}
	// ---- Part B: distributed merge of the splits — CTA c merges outputs [32c, 32c + 32) (one head slice), in split order ----
	if (n_active > 1) {
		const int n_out = n_vkv * G * HD;
		float* s_ms = scratch;                    // [DM_MAX_SPLITS] split maxima of this head
		float* s_num = scratch + DM_MAX_SPLITS;   // [8][32]
		float* s_den = s_num + 8 * 32;            // [8]
		const int ol = threadIdx.x & 31, sg = threadIdx.x >> 5;
		for (int c = (int) blockIdx.x; c * 32 < n_out; c += (int) gridDim.x) {
			const int o = c * 32 + ol;
			const int h = o / HD, d = o % HD, kvh = h / G, g = h % G;
			constexpr int MS = DM_MAX_SPLITS / 8;
			dm_tagged wv[MS], wm[MS], wl[MS];
#pragma unroll
			for (int i = 0; i < MS; i++) {
				const int sp = sg + 8 * i;
				if (sp < n_active) {
					const size_t base = ((size_t) kvh * a.n_splits + sp) * G + g;
					wv[i] = dm_ld(P.tpart + base * HD + d);
					dm_ld2(P.tml + base * 2, wm[i], wl[i]);
				}
			}
			float mv[MS], lv[MS], vv[MS];
#pragma unroll
			for (int i = 0; i < MS; i++) {
				const int sp = sg + 8 * i;
				mv[i] = -CUDART_INF_F; lv[i] = 0.f; vv[i] = 0.f;
				if (sp < n_active) {
					const size_t base = ((size_t) kvh * a.n_splits + sp) * G + g;
					unsigned int spins = 0;
					while (!(dm_fresh(wv[i], tag_out) && dm_fresh(wm[i], tag_out) && dm_fresh(wl[i], tag_out))) {
						if ((++spins & 63u) == 0 && dm_check_abort(ab, spins)) break;
						wv[i] = dm_ld(P.tpart + base * HD + d);
						dm_ld2(P.tml + base * 2, wm[i], wl[i]);
					}
					vv[i] = __uint_as_float((uint32_t) wv[i]);
					mv[i] = __uint_as_float((uint32_t) wm[i]);
					lv[i] = __uint_as_float((uint32_t) wl[i]);
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
				const float r = tn / td;
				dm_st(P.out_t + o, dm_packf(r, tag_out));
				a.out[o] = r;
			}
			dm_bar();
		}
	}
}
template <int G, int KVDIV>
__device__ __forceinline__ void dm_attention_hd(const DmPhase& P, const StepParams& st, float* scratch, int first, int stride, uint32_t tag_in,
                                                uint32_t tag_out, const DmAbort& ab) {
	if (P.HD == 64) dm_attention<64, G, KVDIV>(P, st, scratch, first, stride, tag_in, tag_out, ab);
	else dm_attention<128, G, KVDIV>(P, st, scratch, first, stride, tag_in, tag_out, ab);
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
	__shared__ int s_abort;

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) {
		s_abort = 0;
		for (int s = 0; s < NS; s++) {
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], DM_CW);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < PH_WORDS) reinterpret_cast<uint32_t*>(&s_ph[0])[threadIdx.x] = reinterpret_cast<const uint32_t*>(&mk.phases[0])[threadIdx.x];
	if (threadIdx.x < (int) (sizeof(StepParams) / 4)) reinterpret_cast<uint32_t*>(&s_step)[threadIdx.x] = reinterpret_cast<const uint32_t*>(mk.step)[threadIdx.x];
	if (threadIdx.x < mk.head_dim / 2) s_freq[threadIdx.x] = mk.rope_freq[threadIdx.x];
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
				if (epi == EPI_QKV && ph + 1 < mk.n_phases && mk.phases[ph + 1].kind == DM_ATTN) {
					// the K/V rows the attention phase after this one will walk: this CTA pulls its share into L2 now, so that phase
					// reads them at L2 latency instead of queueing behind the weight stream in DRAM
					const AttnArgs& at = mk.phases[ph + 1].at;
					const unsigned long long kvb = (unsigned long long) s_step.kv_len * at.n_kv_heads * mk.phases[ph + 1].HD * sizeof(__half);
					l2_prefetch_slice(reinterpret_cast<const uint8_t*>(at.k_cache), kvb, (int) blockIdx.x, G);
					l2_prefetch_slice(reinterpret_cast<const uint8_t*>(at.v_cache), kvb, (int) blockIdx.x, G);
				}
				for (int tile = first; tile < n_tiles; tile += G) {
					const int row0 = tile * RC;
					for (int kr = 0; kr < kranges; kr++) {
						if (!dm_mbar_wait(&empty[slot], phase ^ 1, &s_abort)) return;
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
	const int kw = warp % KW, rw = warp / KW;
	const int hA = (lane >> 2) & 1;
	int slot = 0, phase = 0;
	const DmAbort ab = {&s_abort, mk.gbar + 1, mk.err};
	const uint32_t tagbase = s_step.ar_base;
	unsigned long long* tl = (mk.tl && blockIdx.x == 0 && threadIdx.x == 0) ? mk.tl : nullptr;

	for (int ph = 0; ph < mk.n_phases; ph++) {
		const DmPhase& P = s_ph[ph & 1];
		if (tl) tl[4 * ph] = gtime();
		if (mk.tl && threadIdx.x == 0) mk.tl[(size_t) 4 * mk.tl_phases + (size_t) ph * G + blockIdx.x] = gtime();
		// the next phase's descriptor travels while this phase runs (read once per token: DRAM, or L2 thanks to the producer)
		uint32_t next_word = 0;
		const bool has_next = ph + 1 < mk.n_phases && threadIdx.x < PH_WORDS;
		if (has_next) next_word = reinterpret_cast<const uint32_t*>(&mk.phases[ph + 1])[threadIdx.x];
		const uint32_t tag_in = tagbase + (uint32_t) ph + 1u, tag_out = tagbase + (uint32_t) ph + 2u;
		const int first = ((int) blockIdx.x + G - P.tile_off % G) % G;
		if (P.kind == DM_ATTN) {
			float* scratch = reinterpret_cast<float*>(xq_base);
			switch (P.G) {
				case 1: dm_attention_hd<1, 1>(P, s_step, scratch, first, G, tag_in, tag_out, ab); break;
				case 2: dm_attention_hd<2, 1>(P, s_step, scratch, first, G, tag_in, tag_out, ab); break;
				case 4: dm_attention_hd<4, 1>(P, s_step, scratch, first, G, tag_in, tag_out, ab); break;
				case 8: dm_attention_hd<4, 2>(P, s_step, scratch, first, G, tag_in, tag_out, ab); break;
			}
			if (tl) { tl[4 * ph + 1] = tl[4 * ph + 2] = tl[4 * ph]; tl[4 * ph + 3] = gtime(); }
		} else {
			const MatvecArgs& a = P.a;
			const int n = a.n, nu = n / 256, epi = a.epi;
			const int kranges = P.kranges, n_tiles = P.n_tiles;
			dm_tagged* const out_t = P.out_t;
			if (epi == EPI_QKV && blockIdx.x == 0 && s_step.kv_sink > 0) { // attention sinks move on by one position (infer.cpp:416-431)
				const int pairs = a.kv_dim / 2;
				for (int i = threadIdx.x; i < s_step.kv_sink * pairs; i += DM_CW * 32) {
					const int r = i / pairs, p = i % pairs;
					__half2* kp = reinterpret_cast<__half2*>(a.k_cache + (size_t) r * a.kv_dim) + p;
					float2 vv = __half22float2(__ldcg(kp));
					rope_pair(vv.x, vv.y, (2 * p) % a.head_dim, 1, s_freq);
					const __half2 h2 = __floats2half2_rn(vv.x, vv.y);
					*kp = h2;
					dm_st(P.tsink + (size_t) r * pairs + p, dm_pack(*reinterpret_cast<const uint32_t*>(&h2), tag_out));
				}
			}
			if (first < n_tiles) { // a CTA without tiles in this phase reads nothing (and so can never wait for an overwritten tag)
				const XqView xv = xq_view(xq_base, n);
				if (a.norm_w != nullptr) {
					DmNormW nw;
					dm_norm_fetch(a, nw);
					dm_stage_norm(a, P.in_t, tag_in, xv, s_red, reinterpret_cast<float*>(xq_base + mk.xq_cap) - n, nw, ab); // host: xq_cap >= xq_bytes(n) + 4 n
				} else {
					dm_stage_plain(a, P.in_t, tag_in, xv, ab);
				}
				dm_bar();
				if (tl) tl[4 * ph + 1] = tl[4 * ph + 2] = gtime();

				int tcount = 0;
				for (int tile = first; tile < n_tiles; tile += G, tcount++) {
					const int row0 = tile * RC;
					const bool reducer = warp == (tcount % DM_CW);
					// residual: request the old activation early so the epilogue does not sit on a trip to L2
					dm_tagged xold_w = 0;
					if (epi == EPI_RESIDUAL && reducer && lane < RC) xold_w = dm_ld(out_t + row0 + lane);
					float y[R];
#pragma unroll
					for (int r = 0; r < R; r++) y[r] = 0.f;
					for (int kr = 0; kr < kranges; kr++) {
						const int u0 = kr * U;
						const int nb = 8 * min(U, nu - u0); // blocks per row in this stage
						const int b = kw * 32 + lane;
						const bool got = dm_mbar_wait(&full[slot], phase, &s_abort);
						if (got && b < nb) {
							XqBlock xb;
							xq_load(xv, u0 * 8 + b, hA, xb);
							const uint8_t* unit = ring + (size_t) slot * mk.slot_bytes + (size_t) (rw * R) * ROW_STAGE + (size_t) (b >> 3) * UB;
#pragma unroll
							for (int r = 0; r < R; r++) F::block(unit + (size_t) r * ROW_STAGE, b & 7, hA, xb, y[r]);
						}
						__syncwarp();
						if (lane == 0) mbar_arrive(&empty[slot]);
						if (++slot == NS) { slot = 0; phase ^= 1; }
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
								const float xn = __uint_as_float((uint32_t) xold_w) + yv; // infer.cpp:450-452, :492-494
								dm_st(out_t + row, dm_packf(xn, tag_out));
								a.out[row] = xn;
							} else if (epi == EPI_GLU) {
								if ((lane & 1) == 0) { // (W1[o], W3[o]) sit in adjacent rows of a GLU tile
									const float g = a.act == XALM_SILU ? act_silu(yv) : act_gelu(yv);
									const float hv = g * ynext; // infer.cpp:470-488
									dm_st(out_t + (row >> 1), dm_packf(hv, tag_out));
									a.out[row >> 1] = hv;
								}
							} else if (epi == EPI_STORE) {
								if (row < a.d) a.out[row] = yv; // logits (infer.cpp:637)
							} else if ((lane & 1) == 0) { // EPI_QKV: clip -> RoPE -> q, or fp16 K/V into the cache ring (infer.cpp:388-414)
								float v0 = clipf(yv, a.qkv_clip), v1 = clipf(ynext, a.qkv_clip);
								if (row < a.q_dim) {
									rope_pair(v0, v1, row % a.head_dim, s_step.pos, s_freq);
									dm_st2(out_t + row, dm_packf(v0, tag_out), dm_packf(v1, tag_out));
									a.out[row] = v0;
									a.out[row + 1] = v1;
								} else if (row < a.q_dim + a.kv_dim) {
									const int i = row - a.q_dim;
									rope_pair(v0, v1, i % a.head_dim, s_step.pos, s_freq);
									const __half2 h2 = __floats2half2_rn(v0, v1);
									*reinterpret_cast<__half2*>(a.k_cache + (size_t) s_step.kv_pos * a.kv_dim + i) = h2;
									dm_st(P.tkv + i / 2, dm_pack(*reinterpret_cast<const uint32_t*>(&h2), tag_out));
								} else {
									const int i = row - a.q_dim - a.kv_dim;
									const __half2 h2 = __floats2half2_rn(v0, v1);
									*reinterpret_cast<__half2*>(a.v_cache + (size_t) s_step.kv_pos * a.kv_dim + i) = h2;
									dm_st(P.tkv + a.kv_dim / 2 + i / 2, dm_pack(*reinterpret_cast<const uint32_t*>(&h2), tag_out));
								}
							}
						}
					}
				}
			}
			if (tl) tl[4 * ph + 3] = gtime();
		}
		if (has_next) reinterpret_cast<uint32_t*>(&s_ph[(ph + 1) & 1])[threadIdx.x] = next_word;
		dm_bar(); // every warp is done with this phase's staged activations, partial sums and descriptor
	}
	if (threadIdx.x == 0 && s_abort && mk.err) *mk.err = 3u;
}

bool dm_supported_type(int type) { return idp_supported(type); }
size_t dm_xq_bytes(int n) { return xq_bytes(n); }

size_t dm_attn_scratch_bytes(int HD, int G) {
	if (G == 8) G = 4; // served as two passes of four heads
	const int LPR = HD / 8, RPW = 32 / LPR, NGRP = DM_CW * RPW;
	return ((size_t) 3 * NGRP * G + (size_t) DM_CW * G * HD) * sizeof(float);
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
