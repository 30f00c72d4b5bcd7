// ckd_hostmath.h -- host-side scalar helpers used for the per-frame set-up the reference also does on the CPU
// (a handful of values per frame: palettes, rotation cosines, projection tables).  Operation order follows the
// reference op for op so the values handed to the kernels are bit-identical to what the reference's loops see.
// Compile with -ffp-contract=off.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>

namespace ckdh {

constexpr float kPI = 3.1415926535897932384626433832795f;   // Std3DMath-stripped/Math.h:17
constexpr float k2PI = 2.f*kPI;
constexpr float kEpsilon = 1.1920928955078125e-07f;          // FLT_EPSILON
constexpr float kGoldenRatio = 1.61803398875f;
constexpr float kGoldenAngle = 2.39996f;

// x86 cvttss2si: truncate, out-of-range / NaN -> 0x80000000
static inline int x86_cvtt(float f)
{
	if (!(f >= -2147483648.f && f < 2147483648.f))
		return int(0x80000000u);
	return int(f);
}

// gcc on x86-64: unsigned(float) = 64-bit cvttss2si, low 32 bits kept
static inline unsigned x86_f2u(float f)
{
	if (!(f >= -9223372036854775808.f && f < 9223372036854775808.f))
		return 0u;
	return unsigned(uint64_t(int64_t(f)));
}

static inline int ftofp24(float value) { return x86_cvtt(value*256.f); } // util.h:195-197

static inline float fracf(float value) { return value - truncf(value); }  // Math.h:49
static inline float lerpf(float a, float b, float t) { return a + (b-a)*t; } // Math.h:52-56
static inline float stdmax(float a, float b) { return (a < b) ? b : a; }  // std::max<float>
static inline float stdmin(float a, float b) { return (b < a) ? b : a; }  // std::min<float>
static inline float clampf(float mn, float mx, float v) { return stdmax(mn, stdmin(mx, v)); } // Math.h:37-40
static inline float saturatef(float v) { return stdmax(0.f, stdmin(1.f, v)); }                 // Math.h:43-46
static inline int clampi(int mn, int mx, int v) { return std::max(mn, std::min(mx, v)); }      // util.h:205-207

// sincos-lut.h:13-26 (ARRESTED_DEV_LEGACY is defined: lutsinf(a) = lutcosf(a + pi/2))
static inline float lutcosf(const float *lut, float angle)
{
	angle = fabsf(angle);
	angle *= (1.f/k2PI)*2048;
	const int index = x86_cvtt(angle) & 2047;
	return lerpf(lut[index], lut[index+1], fracf(angle));
}
static inline float lutsinf(const float *lut, float angle) { return lutcosf(lut, angle + kPI*0.5f); }

struct vec4 { float x, y, z, w; };

// Shadertoy::MichielPal, shadertoy-util.h:193-198
static inline vec4 MichielPal(const float *lut, float phase)
{
	return { .1f - lutcosf(lut, phase/3.f)/(19.f*0.5f), .1f, .1f + lutcosf(lut, phase/14.f)/4.f, 0.f };
}

// _mm_dp_ps(a, b, 0xff): (a0*b0 + a1*b1) + (a2*b2 + a3*b3)
static inline float dp4(const vec4 &a, const vec4 &b) { return (a.x*b.x + a.y*b.y) + (a.z*b.z + a.w*b.w); }

// Shadertoy::Desaturate, shadertoy-util.h:244-249 (all four lanes, the 4th one matters for __m128 consumers)
static inline vec4 Desaturate(const vec4 &color, float amount)
{
	const vec4 weights = { 0.0722f, 0.7152f, 0.2126f, 0.f };
	const float luma = dp4(weights, color);
	return { color.x + amount*(luma-color.x), color.y + amount*(luma-color.y), color.z + amount*(luma-color.z), color.w + amount*(luma-color.w) };
}

// BoxBlurScale, deprecated/boxblur.h:13-19
static inline float BoxBlurScale(float strength)
{
	if (strength != 0.f)
		strength = clampf(1.f, 100.f, strength)*0.01f;
	return strength;
}

// synth-math-easings.h:190-203
static inline float easeInOutElasticf(float x)
{
	constexpr float c5 = (2.f*kPI)/4.5f;
	return (0.f == x) ? 0.f
		: (1.f == x) ? 1.f
		: (x < 0.5f)
			? -(powf(2.f, 20.f*x - 10.f) * sinf((20.f*x - 11.125f) * c5))*0.5f
			: (powf(2.f, -20.f*x + 10.f) * sinf((20.f*x - 11.125f) * c5))*0.5f + 1.f;
}

} // namespace ckdh
