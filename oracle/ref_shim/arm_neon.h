// arm_neon.h — a scalar/GCC-vector stand-in for the handful of NEON types and intrinsics the reference (jubruckne/Xalm,
// Apple-Silicon only as written: src/types2.h:2, src/infer.cpp:92, src/types.h:13) touches, so that its OWN translation units
// compile unmodified on x86_64 with g++ 13 (oracle/Makefile.ref).  Test infrastructure only: lets the oracle be pinned to the
// reference's real forward pass.  Semantics follow the Arm ACLE definitions element for element; nothing here is tuned.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

typedef float float32_t;
typedef _Float16 float16_t;
typedef double float64_t;
// bf16 storage type with the implicit conversions Apple clang gives __bf16 (widening is a 16-bit shift; narrowing rounds to
// nearest even).  src/types.h:322-335 never relies on more.
struct bfloat16_t {
	uint16_t bits;
	bfloat16_t() = default;
	bfloat16_t(float f) {
		uint32_t u;
		std::memcpy(&u, &f, 4);
		if ((u & 0x7fffffffu) > 0x7f800000u) bits = (uint16_t) ((u >> 16) | 0x40);
		else bits = (uint16_t) ((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
	}
	operator float() const {
		const uint32_t u = (uint32_t) bits << 16;
		float f;
		std::memcpy(&f, &u, 4);
		return f;
	}
};

#define XALM_SHIM_VEC(name, elem, n) typedef elem name __attribute__((vector_size(sizeof(elem) * n)))
XALM_SHIM_VEC(int8x16_t, int8_t, 16);
XALM_SHIM_VEC(uint8x16_t, uint8_t, 16);
XALM_SHIM_VEC(int8x8_t, int8_t, 8);
XALM_SHIM_VEC(uint8x8_t, uint8_t, 8);
XALM_SHIM_VEC(int16x8_t, int16_t, 8);
XALM_SHIM_VEC(uint16x8_t, uint16_t, 8);
XALM_SHIM_VEC(int32x4_t, int32_t, 4);
XALM_SHIM_VEC(uint32x4_t, uint32_t, 4);
XALM_SHIM_VEC(int64x2_t, int64_t, 2);
XALM_SHIM_VEC(uint64x2_t, uint64_t, 2);
XALM_SHIM_VEC(float16x8_t, _Float16, 8);
XALM_SHIM_VEC(float16x4_t, _Float16, 4);
XALM_SHIM_VEC(float32x4_t, float, 4);
XALM_SHIM_VEC(float64x2_t, double, 2);
#undef XALM_SHIM_VEC
struct float32x4x2_t { float32x4_t val[2]; };
struct float32x4x4_t { float32x4_t val[4]; };
struct float16x8x2_t { float16x8_t val[2]; };
struct float16x8x4_t { float16x8_t val[4]; };

#define XALM_SHIM_INLINE static inline __attribute__((always_inline, unused))

// ---- loads / stores ----
#define XALM_SHIM_LDST(sfx, vec, elem)                                                        \
	XALM_SHIM_INLINE vec vld1q_##sfx(const elem* p) { vec v; std::memcpy(&v, p, sizeof v); return v; } \
	XALM_SHIM_INLINE void vst1q_##sfx(elem* p, vec v) { std::memcpy(p, &v, sizeof v); }
XALM_SHIM_LDST(s8, int8x16_t, int8_t)
XALM_SHIM_LDST(u8, uint8x16_t, uint8_t)
XALM_SHIM_LDST(s16, int16x8_t, int16_t)
XALM_SHIM_LDST(u16, uint16x8_t, uint16_t)
XALM_SHIM_LDST(s32, int32x4_t, int32_t)
XALM_SHIM_LDST(u32, uint32x4_t, uint32_t)
XALM_SHIM_LDST(s64, int64x2_t, int64_t)
XALM_SHIM_LDST(u64, uint64x2_t, uint64_t)
XALM_SHIM_LDST(f16, float16x8_t, float16_t)
XALM_SHIM_LDST(f32, float32x4_t, float32_t)
XALM_SHIM_LDST(f64, float64x2_t, float64_t)
#undef XALM_SHIM_LDST
XALM_SHIM_INLINE float32x4x2_t vld1q_f32_x2(const float* p) { return {{vld1q_f32(p), vld1q_f32(p + 4)}}; }
XALM_SHIM_INLINE float32x4x4_t vld1q_f32_x4(const float* p) { return {{vld1q_f32(p), vld1q_f32(p + 4), vld1q_f32(p + 8), vld1q_f32(p + 12)}}; }
XALM_SHIM_INLINE float16x8x2_t vld1q_f16_x2(const float16_t* p) { return {{vld1q_f16(p), vld1q_f16(p + 8)}}; }
XALM_SHIM_INLINE float16x8x4_t vld1q_f16_x4(const float16_t* p) { return {{vld1q_f16(p), vld1q_f16(p + 8), vld1q_f16(p + 16), vld1q_f16(p + 24)}}; }
XALM_SHIM_INLINE void vst1q_f32_x2(float* p, float32x4x2_t v) { vst1q_f32(p, v.val[0]); vst1q_f32(p + 4, v.val[1]); }
XALM_SHIM_INLINE void vst1q_f32_x4(float* p, float32x4x4_t v) { for (int i = 0; i < 4; i++) vst1q_f32(p + 4 * i, v.val[i]); }
XALM_SHIM_INLINE void vst1q_f16_x2(float16_t* p, float16x8x2_t v) { vst1q_f16(p, v.val[0]); vst1q_f16(p + 8, v.val[1]); }
XALM_SHIM_INLINE void vst1q_f16_x4(float16_t* p, float16x8x4_t v) { for (int i = 0; i < 4; i++) vst1q_f16(p + 8 * i, v.val[i]); }

// ---- broadcast ----
XALM_SHIM_INLINE float32x4_t vdupq_n_f32(float v) { return float32x4_t{v, v, v, v}; }
XALM_SHIM_INLINE float16x8_t vdupq_n_f16(float16_t v) { return float16x8_t{v, v, v, v, v, v, v, v}; }
XALM_SHIM_INLINE int32x4_t vdupq_n_s32(int32_t v) { return int32x4_t{v, v, v, v}; }
XALM_SHIM_INLINE uint32x4_t vdupq_n_u32(uint32_t v) { return uint32x4_t{v, v, v, v}; }
XALM_SHIM_INLINE uint8x16_t vdupq_n_u8(uint8_t v) { uint8x16_t r; for (int i = 0; i < 16; i++) r[i] = v; return r; }
XALM_SHIM_INLINE int8x16_t vdupq_n_s8(int8_t v) { int8x16_t r; for (int i = 0; i < 16; i++) r[i] = v; return r; }

// ---- element-wise arithmetic (GCC vector operators are element-wise, wrap-around for integers like NEON) ----
XALM_SHIM_INLINE float32x4_t vaddq_f32(float32x4_t a, float32x4_t b) { return a + b; }
XALM_SHIM_INLINE float32x4_t vmulq_f32(float32x4_t a, float32x4_t b) { return a * b; }
XALM_SHIM_INLINE float32x4_t vfmaq_f32(float32x4_t acc, float32x4_t a, float32x4_t b) { float32x4_t r; for (int i = 0; i < 4; i++) r[i] = std::fmaf(a[i], b[i], acc[i]); return r; }
XALM_SHIM_INLINE float32x4_t vsqrtq_f32(float32x4_t a) { float32x4_t r; for (int i = 0; i < 4; i++) r[i] = std::sqrt(a[i]); return r; }
XALM_SHIM_INLINE float32x4_t vabsq_f32(float32x4_t a) { float32x4_t r; for (int i = 0; i < 4; i++) r[i] = std::fabs(a[i]); return r; }
XALM_SHIM_INLINE float32x4_t vpaddq_f32(float32x4_t a, float32x4_t b) { return float32x4_t{a[0] + a[1], a[2] + a[3], b[0] + b[1], b[2] + b[3]}; }
XALM_SHIM_INLINE float vaddvq_f32(float32x4_t a) { return (a[0] + a[1]) + (a[2] + a[3]); }
XALM_SHIM_INLINE float16x8_t vaddq_f16(float16x8_t a, float16x8_t b) { return a + b; }
XALM_SHIM_INLINE float16x8_t vmulq_f16(float16x8_t a, float16x8_t b) { return a * b; }
XALM_SHIM_INLINE float16x8_t vfmaq_f16(float16x8_t acc, float16x8_t a, float16x8_t b) { float16x8_t r; for (int i = 0; i < 8; i++) r[i] = (float16_t) std::fmaf((float) a[i], (float) b[i], (float) acc[i]); return r; }
XALM_SHIM_INLINE int8x16_t vaddq_s8(int8x16_t a, int8x16_t b) { return a + b; }
XALM_SHIM_INLINE int8x16_t vsubq_s8(int8x16_t a, int8x16_t b) { return a - b; }
XALM_SHIM_INLINE uint8x16_t vaddq_u8(uint8x16_t a, uint8x16_t b) { return a + b; }
XALM_SHIM_INLINE uint8x16_t vsubq_u8(uint8x16_t a, uint8x16_t b) { return a - b; }
XALM_SHIM_INLINE int16x8_t vaddq_s16(int16x8_t a, int16x8_t b) { return a + b; }
XALM_SHIM_INLINE uint16x8_t vaddq_u16(uint16x8_t a, uint16x8_t b) { return a + b; }
XALM_SHIM_INLINE uint8x16_t vandq_u8(uint8x16_t a, uint8x16_t b) { return a & b; }
XALM_SHIM_INLINE uint8x16_t vshrq_n_u8(uint8x16_t a, int n) { return a >> (uint8_t) n; }
XALM_SHIM_INLINE int8x16_t vminq_s8(int8x16_t a, int8x16_t b) { int8x16_t r; for (int i = 0; i < 16; i++) r[i] = a[i] < b[i] ? a[i] : b[i]; return r; }
XALM_SHIM_INLINE int8x16_t vmaxq_s8(int8x16_t a, int8x16_t b) { int8x16_t r; for (int i = 0; i < 16; i++) r[i] = a[i] > b[i] ? a[i] : b[i]; return r; }
XALM_SHIM_INLINE uint8x16_t vminq_u8(uint8x16_t a, uint8x16_t b) { uint8x16_t r; for (int i = 0; i < 16; i++) r[i] = a[i] < b[i] ? a[i] : b[i]; return r; }
XALM_SHIM_INLINE uint8x16_t vmaxq_u8(uint8x16_t a, uint8x16_t b) { uint8x16_t r; for (int i = 0; i < 16; i++) r[i] = a[i] > b[i] ? a[i] : b[i]; return r; }
XALM_SHIM_INLINE int8x16_t vabsq_s8(int8x16_t a) { int8x16_t r; for (int i = 0; i < 16; i++) r[i] = (int8_t) (a[i] < 0 ? -a[i] : a[i]); return r; }

// ---- halves, widening, narrowing ----
XALM_SHIM_INLINE uint8x8_t vget_low_u8(uint8x16_t a) { uint8x8_t r; for (int i = 0; i < 8; i++) r[i] = a[i]; return r; }
XALM_SHIM_INLINE uint8x8_t vget_high_u8(uint8x16_t a) { uint8x8_t r; for (int i = 0; i < 8; i++) r[i] = a[8 + i]; return r; }
XALM_SHIM_INLINE int8x8_t vget_low_s8(int8x16_t a) { int8x8_t r; for (int i = 0; i < 8; i++) r[i] = a[i]; return r; }
XALM_SHIM_INLINE int8x8_t vget_high_s8(int8x16_t a) { int8x8_t r; for (int i = 0; i < 8; i++) r[i] = a[8 + i]; return r; }
XALM_SHIM_INLINE float16x4_t vget_low_f16(float16x8_t a) { return float16x4_t{a[0], a[1], a[2], a[3]}; }
XALM_SHIM_INLINE float16x4_t vget_high_f16(float16x8_t a) { return float16x4_t{a[4], a[5], a[6], a[7]}; }
XALM_SHIM_INLINE float16_t vget_lane_f16(float16x4_t a, int lane) { return a[lane]; }
XALM_SHIM_INLINE float16x8_t vcombine_f16(float16x4_t lo, float16x4_t hi) { return float16x8_t{lo[0], lo[1], lo[2], lo[3], hi[0], hi[1], hi[2], hi[3]}; }
XALM_SHIM_INLINE uint16x8_t vmovl_u8(uint8x8_t a) { uint16x8_t r; for (int i = 0; i < 8; i++) r[i] = a[i]; return r; }
XALM_SHIM_INLINE uint16x8_t vmull_u8(uint8x8_t a, uint8x8_t b) { uint16x8_t r; for (int i = 0; i < 8; i++) r[i] = (uint16_t) ((uint16_t) a[i] * (uint16_t) b[i]); return r; }
XALM_SHIM_INLINE int16x8_t vmull_s8(int8x8_t a, int8x8_t b) { int16x8_t r; for (int i = 0; i < 8; i++) r[i] = (int16_t) ((int16_t) a[i] * (int16_t) b[i]); return r; }
XALM_SHIM_INLINE uint8x8_t vmovn_u16(uint16x8_t a) { uint8x8_t r; for (int i = 0; i < 8; i++) r[i] = (uint8_t) a[i]; return r; }
XALM_SHIM_INLINE int8x8_t vmovn_s16(int16x8_t a) { int8x8_t r; for (int i = 0; i < 8; i++) r[i] = (int8_t) a[i]; return r; }
XALM_SHIM_INLINE uint8x16_t vcombine_u8(uint8x8_t lo, uint8x8_t hi) { uint8x16_t r; for (int i = 0; i < 8; i++) { r[i] = lo[i]; r[8 + i] = hi[i]; } return r; }
XALM_SHIM_INLINE int8x16_t vcombine_s8(int8x8_t lo, int8x8_t hi) { int8x16_t r; for (int i = 0; i < 8; i++) { r[i] = lo[i]; r[8 + i] = hi[i]; } return r; }
XALM_SHIM_INLINE float32x4_t vcvt_f32_f16(float16x4_t a) { return float32x4_t{(float) a[0], (float) a[1], (float) a[2], (float) a[3]}; }
XALM_SHIM_INLINE float16x4_t vcvt_f16_f32(float32x4_t a) { return float16x4_t{(float16_t) a[0], (float16_t) a[1], (float16_t) a[2], (float16_t) a[3]}; }
XALM_SHIM_INLINE float16x8_t vcvtq_f16_u16(uint16x8_t a) { float16x8_t r; for (int i = 0; i < 8; i++) r[i] = (float16_t) a[i]; return r; }

// ---- pairwise / across / dot ----
XALM_SHIM_INLINE int32x4_t vpaddlq_s16(int16x8_t a) { int32x4_t r; for (int i = 0; i < 4; i++) r[i] = (int32_t) a[2 * i] + (int32_t) a[2 * i + 1]; return r; }
XALM_SHIM_INLINE uint32x4_t vpaddlq_u16(uint16x8_t a) { uint32x4_t r; for (int i = 0; i < 4; i++) r[i] = (uint32_t) a[2 * i] + (uint32_t) a[2 * i + 1]; return r; }
XALM_SHIM_INLINE int32_t vaddvq_s32(int32x4_t a) { return a[0] + a[1] + a[2] + a[3]; }
XALM_SHIM_INLINE uint32_t vaddvq_u32(uint32x4_t a) { return a[0] + a[1] + a[2] + a[3]; }
XALM_SHIM_INLINE int32x4_t vdotq_s32(int32x4_t acc, int8x16_t a, int8x16_t b) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc[i] += (int32_t) a[4 * i + j] * (int32_t) b[4 * i + j]; return acc; }
XALM_SHIM_INLINE uint32x4_t vdotq_u32(uint32x4_t acc, uint8x16_t a, uint8x16_t b) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc[i] += (uint32_t) a[4 * i + j] * (uint32_t) b[4 * i + j]; return acc; }
