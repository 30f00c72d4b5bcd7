// ckd_math.cuh -- device math layer: x86/SSE arithmetic semantics restated for sm_100a.
//
// The reference is an SSE4.1 + glibc program; its pixels depend on x86 conversion rules, SSE min/max NaN rules,
// the Cephes-derived log_ps/exp_ps polynomials, the CPU's RSQRTPS approximation and on the absence of FMA contraction
// (SURVEY.md appendix A).  Everything here is written so that, compiled with -fmad=false (and the default
// -prec-div=true -prec-sqrt=true -ftz=false), each function returns the same bits as its x86 counterpart.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ckd {

constexpr float kPI = 3.1415926535897932384626433832795f;   // Std3DMath-stripped/Math.h:17
constexpr float k2PI = 2.f*kPI;
constexpr float kEpsilon = 1.1920928955078125e-07f;          // FLT_EPSILON (Math.h:20)
constexpr float kGoldenRatio = 1.61803398875f;
constexpr float kGoldenAngle = 2.39996f;

// ---- conversions --------------------------------------------------------------------------------------------------

// cvttss2si (C cast on x86): truncate; out of range or NaN -> 0x80000000
__device__ __forceinline__ int cvtt_x86(float f)
{
	const int r = __float2int_rz(f);
	return (f >= -2147483648.f && f < 2147483648.f) ? r : int(0x80000000u);
}

// cvtps2dq (_mm_cvtps_epi32): round to nearest even; out of range or NaN -> 0x80000000
__device__ __forceinline__ int cvtn_x86(float f)
{
	const int r = __float2int_rn(f);
	return (f >= -2147483648.f && f < 2147483648.f) ? r : int(0x80000000u);
}

// gcc x86-64 'unsigned(float)': 64-bit cvttss2si, low 32 bits
__device__ __forceinline__ unsigned f2u_x86(float f)
{
	if (!(f >= -9223372036854775808.f && f < 9223372036854775808.f))
		return 0u;
	return unsigned(__float2ll_rz(f));
}

__device__ __forceinline__ int ftofp24(float value) { return cvtt_x86(value*256.f); } // util.h:195-197

// ---- scalar helpers with the reference's NaN behaviour ------------------------------------------------------------

__device__ __forceinline__ float stdmax(float a, float b) { return (a < b) ? b : a; } // std::max<float>(a, b)
__device__ __forceinline__ float stdmin(float a, float b) { return (b < a) ? b : a; } // std::min<float>(a, b)
__device__ __forceinline__ float sse_min(float a, float b) { return (a < b) ? a : b; } // MINPS: 2nd operand on NaN / equal
__device__ __forceinline__ float sse_max(float a, float b) { return (a > b) ? a : b; } // MAXPS
__device__ __forceinline__ float clampf(float mn, float mx, float v) { return stdmax(mn, stdmin(mx, v)); } // Math.h:37-40
__device__ __forceinline__ float fracf(float v) { return v - truncf(v); }               // Math.h:49
__device__ __forceinline__ float lerpf(float a, float b, float t) { return a + (b-a)*t; } // Math.h:52-56
__device__ __forceinline__ float smoothstepf(float a, float b, float t)                 // Math.h:59-63
{
	t = t*t * (3.f - 2.f*t);
	return lerpf(a, b, t);
}

// ---- interpolated cosine LUT (sincos-lut.h:13-26) -----------------------------------------------------------------
// The 2049-entry table is staged in shared memory as 2048 (LUT[i], LUT[i+1]-LUT[i]) pairs: one 8-byte LDS per call and the
// difference (the same IEEE subtraction the reference performs per call) is paid once at start-up.

// The table is addressed through its 32-bit shared-window address (CosLut) and read with ld.shared: a generic float2* makes
// the compiler rebuild the shared base (S2UR + UMOV + ULEA) next to every lookup, and these kernels are issue bound.
typedef unsigned CosLut;

__device__ __forceinline__ float lutcosf(CosLut lut, float angle)
{
	angle = fabsf(angle);
	angle *= (1.f/k2PI)*2048; // constant folds in float exactly like the reference's (1.f/k2PI)*kCosTabSize
	// int(angle) & 2047 with cvttss2si semantics: angle is >= 0 or NaN here, so only the upper bound can overflow
	// (-> 0x80000000, index 0); NaN converts to 0 on both machines.  The unsigned conversion is exact below 2^32 and saturates
	// above: bit 31 of it is set exactly when cvttss2si overflows, and its sign smear clears the index -- one shift instead of
	// a compare and a select.  (An F2I-free variant built on the 2^23 rounding trick was measured slower: these kernels are
	// issue bound, not XU bound -- profiles/r01_notes.md.)
	// Written as PRMT (byte 3's sign replicated into every byte) + one three-input LOP3 on the pre-shifted value: left to
	// itself ptxas turns the C expression into ISETP + SEL next to the shift and the mask, one issue slot more per lookup.
	const unsigned u = __float2uint_rz(angle);
	unsigned smear, offset;
	asm("prmt.b32 %0, %1, 0, 0xbbbb;" : "=r"(smear) : "r"(u));
	asm("lop3.b32 %0, %1, 0x3ff8, %2, 0x40;" : "=r"(offset) : "r"(u << 3), "r"(smear)); // a & b & ~c = (u*8) & (2047*8) & ~smear
	float2 pair;                               // (LUT[i], LUT[i+1] - LUT[i])
	asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pair.x), "=f"(pair.y) : "r"(lut + offset));
	return pair.x + pair.y*(angle - truncf(angle)); // lerpf(a, b, t) = a + (b-a)*t, Math.h:52-56
}

// The same lookup without a single conversion instruction, for the march loops of the raymarchers, where F2I + FRND keep
// the XU pipe as busy as the issue port (ncu: XU 70 %, issue 85 %; profiles/r01_notes.md).  For 0 <= angle < 2^23,
// t = angle + 2^23 rounded TOWARDS ZERO is exactly 2^23 + trunc(angle): the low mantissa bits of t are the integer part
// (-> table offset) and t - 2^23 is truncf(angle) (-> fraction), all on the FMA/ALU pipes.  NaN and +inf come out as NaN
// from either variant (NaN fraction times a table entry).  For FINITE scaled angles >= 2^23 (|x| >= 25736 rad) the trick
// does not hold -- the reference then indexes with the low bits of the integer and a zero fraction -- so a kernel may be
// built on this variant only where the host has PROVED every angle of the frame smaller (LutRangeProof in
// ckd_raymarch.cu: effects whose distance functions are bounded, checked against the frame's parameters); every other
// frame runs the exact lutcosf above.
struct CosLutFast { unsigned handle; };

__device__ __forceinline__ float lutcosf(const CosLutFast &lut, float angle)
{
	angle = fabsf(angle);
	angle *= (1.f/k2PI)*2048;
	const float t = __fadd_rz(angle, 8388608.f);
	const unsigned offset = (__float_as_uint(t) << 3) & 0x3ff8u;
	float2 pair;
	asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pair.x), "=f"(pair.y) : "r"(lut.handle + offset));
	return pair.x + pair.y*(angle - (t - 8388608.f));
}

// ARRESTED_DEV_LEGACY (main.h:9): lutsinf(a) = lutcosf(a + pi/2)
template <class Lut> __device__ __forceinline__ float lutsinf(const Lut &lut, float angle)
{
	return lutcosf(lut, angle + kPI*0.5f);
}

// the same on the table where it lies in global memory (TapeWarp32: two lookups per pixel, not worth staging)
__device__ __forceinline__ float lutcosf(const float2 *__restrict__ lut2, float angle)
{
	angle = fabsf(angle);
	angle *= (1.f/k2PI)*2048;
	const unsigned u = __float2uint_rz(angle);
	const float2 pair = __ldg(lut2 + (u & 2047u & ~unsigned(int(u) >> 31)));
	return pair.x + pair.y*(angle - truncf(angle));
}
__device__ __forceinline__ float lutsinf(const float2 *__restrict__ lut2, float angle)
{
	return lutcosf(lut2, angle + kPI*0.5f);
}

// copies the table into shared memory (all threads of the CTA), synchronises, and returns its handle.  The handle is produced
// by a volatile asm *after* the barrier, so no lookup (a pure asm that depends on it) can be scheduled above the barrier.
__device__ __forceinline__ CosLut stage_cos_lut(float2 *s_lut2, const float2 *__restrict__ g_lut2)
{
	for (int i = threadIdx.y*blockDim.x + threadIdx.x; i < 2048; i += blockDim.x*blockDim.y)
		s_lut2[i] = g_lut2[i];
	__syncthreads();
	CosLut lut;
	asm volatile("mov.u32 %0, %1;" : "=r"(lut) : "r"(unsigned(__cvta_generic_to_shared(s_lut2))) : "memory");
	return lut;
}

// ---- RSQRTPS emulation (shadertoy-util.h:89,260,265) --------------------------------------------------------------
// table[parity][mantissa >> log2Bin] = result bits for x = 2^(parity-1)*1.m captured from the host CPU at start-up.
// rsqrt(x * 4^k) = rsqrt(x) * 2^-k is exact for the hardware approximation, so other exponents only shift the exponent.

struct RsqrtTab { const uint32_t *__restrict__ table; int log2Bin; };

__device__ __forceinline__ float rsqrt_x86(const RsqrtTab tab, float x)
{
	const uint32_t bits = __float_as_uint(x);
	const uint32_t exponent = (bits >> 23) & 0xff;
	const uint32_t mantissa = bits & 0x7fffff;

	if (exponent == 0xff)
	{
		if (mantissa) return __uint_as_float(bits | 0x00400000u);    // NaN -> quiet NaN
		return (bits >> 31) ? __uint_as_float(0xffc00000u) : 0.f;     // -inf -> NaN, +inf -> +0
	}
	if (exponent == 0)                                               // zero / denormal -> signed infinity (denormals are treated as zero)
		return __uint_as_float((bits & 0x80000000u) | 0x7f800000u);
	if (bits >> 31)
		return __uint_as_float(0xffc00000u);                         // negative -> default NaN

	const int k = (int(exponent) - 126) >> 1;                       // floor
	const uint32_t parity = (exponent - 126u) & 1u;
	const uint32_t entry = __ldg(tab.table + ((size_t(parity) << (23 - tab.log2Bin)) + (mantissa >> tab.log2Bin)));
	return __uint_as_float(entry - (uint32_t(k) << 23));
}

// ---- Cephes log / exp as restated by sse_mathfun.h (3rdparty/sse_mathfun.h:128-214, 230-306), one lane ------------

__device__ __forceinline__ float log_ps1(float x)
{
	const bool invalid = (x <= 0.f); // cmple: false for NaN

	x = sse_max(x, __uint_as_float(0x00800000u)); // cut off denormalized stuff (NaN -> min_norm_pos)

	int emm0 = int(__float_as_uint(x) >> 23);
	x = __uint_as_float((__float_as_uint(x) & ~0x7f800000u) | 0x3f000000u); // keep mantissa, or 0.5

	emm0 -= 0x7f;
	float e = float(emm0);
	e = e + 1.f;

	const bool mask = (x < 0.707106781186547524f);
	const float tmp = mask ? x : 0.f;
	x = x - 1.f;
	e = e - (mask ? 1.f : 0.f);
	x = x + tmp;

	const float z = x*x;

	float y = 7.0376836292E-2f;
	y = y*x; y = y + -1.1514610310E-1f;
	y = y*x; y = y + 1.1676998740E-1f;
	y = y*x; y = y + -1.2420140846E-1f;
	y = y*x; y = y + 1.4249322787E-1f;
	y = y*x; y = y + -1.6668057665E-1f;
	y = y*x; y = y + 2.0000714765E-1f;
	y = y*x; y = y + -2.4999993993E-1f;
	y = y*x; y = y + 3.3333331174E-1f;
	y = y*x;

	y = y*z;

	float t = e * -2.12194440e-4f;
	y = y + t;

	t = z * 0.5f;
	y = y - t;

	t = e * 0.693359375f;
	x = x + y;
	x = x + t;
	return invalid ? __uint_as_float(0xffffffffu) : x; // or with all-ones mask
}

__device__ __forceinline__ float exp_ps1(float x)
{
	x = sse_min(x, 88.3762626647949f);  // NaN -> exp_hi
	x = sse_max(x, -88.3762626647949f);

	float fx = x * 1.44269504088896341f;
	fx = fx + 0.5f;

	// floorf via truncation (cvttps2dq) + correction
	const int emm0 = cvtt_x86(fx);
	float tmp = float(emm0);
	const float mask = (tmp > fx) ? 1.f : 0.f;
	fx = tmp - mask;

	tmp = fx * 0.693359375f;
	float z = fx * -2.12194440e-4f;
	x = x - tmp;
	x = x - z;

	z = x*x;

	float y = 1.9875691500E-4f;
	y = y*x; y = y + 1.3981999507E-3f;
	y = y*x; y = y + 8.3334519073E-3f;
	y = y*x; y = y + 4.1665795894E-2f;
	y = y*x; y = y + 1.6666665459E-1f;
	y = y*x; y = y + 5.0000001201E-1f;
	y = y*z;
	y = y + x;
	y = y + 1.f;

	// build 2^n
	int n = cvtt_x86(fx);
	n = n + 0x7f;
	n = int(uint32_t(n) << 23);
	return y * __int_as_float(n);
}

// Shadertoy::GammaAdj, shadertoy-util.h:186-190 (one lane)
__device__ __forceinline__ float gamma_adj1(float c, float gamma) { return exp_ps1(gamma * log_ps1(c)); }

// ---- pixel packing ------------------------------------------------------------------------------------------------

// one lane of ToPixel4 (shadertoy-util.h:136-146): max(0, cvtps2dq(255*c)) -> packus32 -> packus16
__device__ __forceinline__ uint32_t to_chan(float c)
{
	int v = cvtn_x86(255.f * c);
	v = max(0, v);
	v = min(v, 65535);  // packus_epi32 (signed 32 -> unsigned 16, v >= 0 here)
	return (v > 32767) ? 0u : uint32_t(min(v, 255)); // packus_epi16 reads the word as signed: >= 0x8000 -> 0
}

__device__ __forceinline__ uint32_t to_pixel(float b, float g, float r, float a)
{
	return to_chan(b) | (to_chan(g) << 8) | (to_chan(r) << 16) | (to_chan(a) << 24);
}

// one lane of ToPixel4_NoConv (shadertoy-util.h:148-157): cvtps2dq -> packus32 -> packus16
__device__ __forceinline__ uint32_t to_chan_noconv(float c)
{
	int v = cvtn_x86(c);
	v = max(0, min(v, 65535));
	return (v > 32767) ? 0u : uint32_t(min(v, 255));
}

// ---- libm calls of the reference (host glibc): evaluated in double and rounded once -------------------------------
// glibc's powf/expf are computed in double internally and are correctly rounded in all but rare cases (measured here:
// (float)exp(double) differs from expf in 0.056 % and (float)pow from powf in 0.065 % of random arguments, always by
// 1 ulp -- far below the 8-bit output quantisation, DESIGN.md "tolerance").

__device__ __forceinline__ float powf_ref(float x, float y) { return float(pow(double(x), double(y))); }
__device__ __forceinline__ float expf_ref(float x) { return float(exp(double(x))); }

// atan2f / atanf as computed by the reference's libm (glibc 2.39 still ships the fdlibm single-precision algorithm:
// table-driven argument reduction + an odd/even split polynomial, all in float).  Restated op for op; verified against
// the host's atan2f on 4e7 random inputs (0 mismatches, tools/check_libm_port.c).
__device__ __forceinline__ float atanf_fdlibm(float x)
{
	const float atanhi[4] = { 4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f };
	const float atanlo[4] = { 5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f };
	const int hx = __float_as_int(x);
	const int ix = hx & 0x7fffffff;
	int id;
	if (ix >= 0x4c000000) // |x| >= 2^25
	{
		if (ix > 0x7f800000) return x+x;
		return (hx > 0) ? atanhi[3]+atanlo[3] : -atanhi[3]-atanlo[3];
	}
	if (ix < 0x3ee00000) // |x| < 0.4375
	{
		if (ix < 0x31000000) return x; // |x| < 2^-29
		id = -1;
	}
	else
	{
		x = fabsf(x);
		if (ix < 0x3f980000) // |x| < 1.1875
		{
			if (ix < 0x3f300000) { id = 0; x = (2.0f*x-1.0f)/(2.0f+x); }
			else                 { id = 1; x = (x-1.0f)/(x+1.0f); }
		}
		else
		{
			if (ix < 0x401c0000) { id = 2; x = (x-1.5f)/(1.0f+1.5f*x); }
			else                 { id = 3; x = -1.0f/x; }
		}
	}
	const float z = x*x;
	const float w = z*z;
	const float s1 = z*(3.3333334327e-01f+w*(1.4285714924e-01f+w*(9.0908870101e-02f+w*(6.6610731184e-02f+w*(4.9768779427e-02f+w*1.6285819933e-02f)))));
	const float s2 = w*(-2.0000000298e-01f+w*(-1.1111110449e-01f+w*(-7.6918758452e-02f+w*(-5.8335702866e-02f+w*-3.6531571299e-02f))));
	if (id < 0) return x - x*(s1+s2);
	const float r = atanhi[id] - ((x*(s1+s2) - atanlo[id]) - x);
	return (hx < 0) ? -r : r;
}

__device__ __forceinline__ float atan2f_ref(float y, float x)
{
	const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
	const int hx = __float_as_int(x), hy = __float_as_int(y);
	const int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
	if (ix > 0x7f800000 || iy > 0x7f800000) return x+y;
	if (hx == 0x3f800000) return atanf_fdlibm(y);
	const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
	if (iy == 0)
	{
		if (m < 2) return y;
		return (m == 2) ? pi+tiny : -pi-tiny;
	}
	if (ix == 0) return (hy < 0) ? -pi_o_2-tiny : pi_o_2+tiny;
	if (ix == 0x7f800000)
	{
		if (iy == 0x7f800000)
		{
			switch (m) { case 0: return pi_o_4+tiny; case 1: return -pi_o_4-tiny; case 2: return 3.0f*pi_o_4+tiny; default: return -3.0f*pi_o_4-tiny; }
		}
		switch (m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return pi+tiny; default: return -pi-tiny; }
	}
	if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2-tiny : pi_o_2+tiny;
	const int k = (iy-ix) >> 23;
	float z;
	if (k > 60) z = pi_o_2+0.5f*pi_lo;
	else if (hx < 0 && k < -60) z = 0.0f;
	else z = atanf_fdlibm(fabsf(y/x));
	switch (m)
	{
	case 0: return z;
	case 1: return __int_as_float(__float_as_int(z) ^ int(0x80000000u));
	case 2: return pi-(z-pi_lo);
	default: return (z-pi_lo)-pi;
	}
}

// Shadertoy::ExpFog, shadertoy-util.h:237-241
__device__ __forceinline__ float exp_fog(float distance, float scale)
{
	return 1.f - (expf_ref(-scale*distance*distance*distance));
}

// Q3_rsqrtf<2>, q3-rsqrt.h:22-42
__device__ __forceinline__ float q3_rsqrtf2(float x)
{
	const float half = 0.5f*x;
	int iX = __float_as_int(x);
	iX = 0x5f3759df - (iX >> 1);
	x = __int_as_float(iX);
	x = x*(1.5f - half*x*x);
	x = x*(1.5f - half*x*x);
	return x;
}

// ---- 3-vectors with the reference's evaluation order ---------------------------------------------------------------

struct vec3 { float x, y, z; };

// _mm_dp_ps(v, v, 0xff) with the Vector3 padding lane = 0: (x*x + y*y) + (z*z + 0*0)
__device__ __forceinline__ float dp_ps3(const vec3 &a, const vec3 &b) { return (a.x*b.x + a.y*b.y) + (a.z*b.z + 0.f); }
// Vector3::Dot, Vector3.h:21-24
__device__ __forceinline__ float dot3(const vec3 &a, const vec3 &b) { return a.x*b.x + a.y*b.y + a.z*b.z; }

// Shadertoy::vNorm4 / vFastNorm3, shadertoy-util.h:78-101
__device__ __forceinline__ void fast_norm3(const RsqrtTab tab, vec3 &v)
{
	const float oneOverLen = rsqrt_x86(tab, dp_ps3(v, v));
	v.x *= oneOverLen; v.y *= oneOverLen; v.z *= oneOverLen;
}

// Shadertoy::vFastLen3, shadertoy-util.h:69-76
__device__ __forceinline__ float fast_len3(const vec3 &v) { return sqrtf(dp_ps3(v, v)); }

// Shadertoy::rotX/rotY/rotZ, shadertoy-util.h:31-59
template <class Lut> __device__ __forceinline__ void rotX(const Lut &lut, float angle, float &Y, float &Z)
{
	const float cosine = lutcosf(lut, angle), sine = lutsinf(lut, angle);
	const float rY = cosine*Y + -sine*Z;
	const float rZ = sine*Y + cosine*Z;
	Y = rY; Z = rZ;
}
template <class Lut> __device__ __forceinline__ void rotY(const Lut &lut, float angle, float &X, float &Z)
{
	const float cosine = lutcosf(lut, angle), sine = lutsinf(lut, angle);
	const float rX = cosine*X + sine*Z;
	const float rZ = -sine*X + cosine*Z;
	X = rX; Z = rZ;
}
template <class Lut> __device__ __forceinline__ void rotZ(const Lut &lut, float angle, float &X, float &Y)
{
	const float cosine = lutcosf(lut, angle), sine = lutsinf(lut, angle);
	const float rX = cosine*X + sine*Y;
	const float rY = -sine*X + cosine*Y;
	X = rX; Y = rY;
}

// ---- integer pixel helpers ------------------------------------------------------------------------------------------

// 8-bit lerp used by every bilinear sampler (bilinear.h:34-125): ((a<<8) + (b-a)*f) >> 8, two channels per 32-bit word.
// a*(256-f) + b*f is the same non-negative 16-bit value, so the packed form is exact.
// lerp8x2_raw leaves the two results in bytes 1 and 3 of the word; the byte permutes below pick them up (one PRMT where a shift
// and a mask would be two ALU instructions: the polar remap and the voxel casters are bound by that pipe / by issue slots).
__device__ __forceinline__ uint32_t lerp8x2_raw(uint32_t a, uint32_t b, uint32_t f) { return a*(256u - f) + b*f; }
__device__ __forceinline__ uint32_t odd_bytes(uint32_t x) { return __byte_perm(x, 0u, 0x4341); }            // (x >> 8) & 0x00ff00ff
__device__ __forceinline__ uint32_t lerp8x2(uint32_t a, uint32_t b, uint32_t f) { return odd_bytes(lerp8x2_raw(a, b, f)); }
// the same lerp on a whole packed pixel, every channel by the same f (Mix32-style blends, the A-variant of the polar blit)
__device__ __forceinline__ uint32_t lerp8x4(uint32_t a, uint32_t b, uint32_t f)
{
	const uint32_t rb = lerp8x2_raw(a & 0x00ff00ffu, b & 0x00ff00ffu, f), ag = lerp8x2_raw(odd_bytes(a), odd_bytes(b), f);
	return __byte_perm(rb, ag, 0x7351); // bytes: rb.1, ag.1, rb.3, ag.3
}

// bsamp32_16 / bsamp32_32 (bilinear.h:58-125) on packed ARGB: returns packed 8-bit channels
__device__ __forceinline__ uint32_t bilerp_argb(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t fu, uint32_t fv)
{
	const uint32_t rb01 = lerp8x2(s0 & 0x00ff00ffu, s1 & 0x00ff00ffu, fu);
	const uint32_t ag01 = lerp8x2(odd_bytes(s0), odd_bytes(s1), fu);
	const uint32_t rb23 = lerp8x2(s2 & 0x00ff00ffu, s3 & 0x00ff00ffu, fu);
	const uint32_t ag23 = lerp8x2(odd_bytes(s2), odd_bytes(s3), fu);
	return __byte_perm(lerp8x2_raw(rb01, rb23, fv), lerp8x2_raw(ag01, ag23, fv), 0x7351);
}

// bsamp8 (bilinear.h:34-53); signed arithmetic shifts as in the reference
__device__ __forceinline__ unsigned bilerp_u8(int s0, int s1, int s2, int s3, int fu, int fv)
{
	const int s01 = ((s0 << 8) + (s1 - s0)*fu) >> 8;
	const int s23 = ((s2 << 8) + (s3 - s2)*fu) >> 8;
	return unsigned(((s01 << 8) + (s23 - s01)*fv) >> 8);
}

// per-byte helpers (SIMD-in-a-word; compile to native video / LOP3 / PRMT sequences)
__device__ __forceinline__ uint32_t avg_u8x4(uint32_t a, uint32_t b) { return __vavgu4(a, b); }   // pavgb: (a+b+1)>>1
__device__ __forceinline__ uint32_t adds_u8x4(uint32_t a, uint32_t b) { return __vaddus4(a, b); } // packus(add_epi16)
__device__ __forceinline__ uint32_t subs_u8x4(uint32_t a, uint32_t b) { return __vsubus4(a, b); } // packus(sub_epi16)

} // namespace ckd
