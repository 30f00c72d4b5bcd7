/*
 * oracle/ckd_oracle.c -- TEST INFRASTRUCTURE ONLY: plain, sequential C restatement of the reference's per-pixel hot path.
 *
 * Nothing under cookiedough_b200/ may include, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may.  It exists so parity can be checked on machines where the reference itself
 * (oracle/_ref, built by oracle/build_ref.py) is not available, and as a second, independently written reading of
 * the reference's arithmetic: scalar loops in the reference's own order (one ray, one step, one pixel at a time),
 * no SIMD, no threads.
 *
 * PINNED: tests/test_oracle_port.py checks every function below against the committed golden fixtures
 * (tests/golden/, generated from the compiled reference by tests/golden/make_golden.py) and, when oracle/_ref is
 * present, against the live reference.
 *
 * Each function cites the reference lines it follows (paths relative to the reference's code/ directory).
 * Host libm (powf/expf/atan2f/sinf/cosf) is called exactly where the reference calls it; the CPU's RSQRTPS is
 * passed in as a table so results do not depend on the machine running the oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KPI 3.1415926535897932384626433832795f
#define K2PI (2.f*KPI)
#define KEPSILON 1.1920928955078125e-07f
#define KGOLDENRATIO 1.61803398875f
#define KGOLDENANGLE 2.39996f

/* ------------------------------------------------------------------------------------------------------------------
 * x86 semantics helpers
 * ---------------------------------------------------------------------------------------------------------------- */

static int cvtt(float f) { return (f >= -2147483648.f && f < 2147483648.f) ? (int)f : (int)0x80000000u; }      /* cvttss2si */
static int cvtn(float f) { return (f >= -2147483648.f && f < 2147483648.f) ? (int)lrintf(f) : (int)0x80000000u; } /* cvtps2dq, RNE */
static unsigned f2u(float f) { return (f >= -9223372036854775808.f && f < 9223372036854775808.f) ? (unsigned)(uint64_t)(int64_t)f : 0u; }
static int ftofp24(float v) { return cvtt(v*256.f); }                    /* util.h:195-197 */
static float stdmaxf(float a, float b) { return (a < b) ? b : a; }       /* std::max<float> */
static float stdminf(float a, float b) { return (b < a) ? b : a; }       /* std::min<float> */
static float ssemin(float a, float b) { return (a < b) ? a : b; }        /* MINPS */
static float ssemax(float a, float b) { return (a > b) ? a : b; }        /* MAXPS */
static float clampf_(float mn, float mx, float v) { return stdmaxf(mn, stdminf(mx, v)); }   /* Math.h:37-40 */
static float saturatef_(float v) { return stdmaxf(0.f, stdminf(1.f, v)); }                  /* Math.h:43-46 */
static float fracf_(float v) { return v - truncf(v); }                   /* Math.h:49 */
static float lerpf_(float a, float b, float t) { return a + (b-a)*t; }   /* Math.h:52-56 */
static float smoothstepf_(float a, float b, float t) { t = t*t*(3.f - 2.f*t); return lerpf_(a, b, t); } /* Math.h:59-63 */
static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ------------------------------------------------------------------------------------------------------------------
 * shared state set by the caller: cosine LUT, RSQRTPS table, resolution
 * ---------------------------------------------------------------------------------------------------------------- */

static float s_cosLUT[2049];
static const uint32_t *s_rsqrtTab;   /* 2 x 1024 entries (parity, top 10 mantissa bits) */
static int s_resX = 1280, s_resY = 720, s_fxX = 644, s_fxY = 364;
static float s_aspect = 0.5625f, s_oneOverAspect = 1.f/0.5625f;

/* CalculateCosLUT, sincos-lut.cpp:9-16 */
void orc_init(int resX, int resY, const uint32_t *rsqrtTable2048)
{
	for (unsigned i = 0; i < 2048; ++i)
		s_cosLUT[i] = cosf((float)i*(K2PI/2048));
	s_cosLUT[2048] = s_cosLUT[0];
	s_rsqrtTab = rsqrtTable2048;
	s_resX = resX; s_resY = resY;
	s_fxX = resX/2 + 4; s_fxY = resY/2 + 4;           /* fx-blitter.h:16-17 */
	s_aspect = (float)resY/(float)resX;               /* main.h:43-44 */
	s_oneOverAspect = 1.f/s_aspect;
}

const float *orc_cos_lut(void) { return s_cosLUT; }

/* lutcosf / lutsinf, sincos-lut.h:13-26 (ARRESTED_DEV_LEGACY) */
float orc_lutcosf(float angle)
{
	angle = fabsf(angle);
	angle *= (1.f/K2PI)*2048;
	const int index = cvtt(angle) & 2047;
	return lerpf_(s_cosLUT[index], s_cosLUT[index+1], fracf_(angle));
}
float orc_lutsinf(float angle) { return orc_lutcosf(angle + KPI*0.5f); }

/* _mm_rsqrt_ps through the captured table (shadertoy-util.h:89) */
float orc_rsqrt(float x)
{
	const uint32_t bits = fbits(x), e = (bits >> 23) & 0xff, m = bits & 0x7fffff;
	if (e == 0xff) return m ? bitsf(bits | 0x00400000u) : ((bits >> 31) ? bitsf(0xffc00000u) : 0.f);
	if (e == 0) return bitsf((bits & 0x80000000u) | 0x7f800000u);
	if (bits >> 31) return bitsf(0xffc00000u);
	const int k = ((int)e - 126) >> 1;
	const uint32_t parity = (e - 126u) & 1u;
	return bitsf(s_rsqrtTab[parity*1024 + (m >> 13)] - ((uint32_t)k << 23));
}

/* log_ps / exp_ps, 3rdparty/sse_mathfun.h:128-214, 230-306 (one lane) */
float orc_log_ps(float x)
{
	const int invalid = (x <= 0.f);
	x = ssemax(x, bitsf(0x00800000u));
	int emm0 = (int)(fbits(x) >> 23);
	x = bitsf((fbits(x) & ~0x7f800000u) | 0x3f000000u);
	emm0 -= 0x7f;
	float e = (float)emm0;
	e = e + 1.f;
	const int mask = (x < 0.707106781186547524f);
	const float tmp0 = mask ? x : 0.f;
	x = x - 1.f;
	e = e - (mask ? 1.f : 0.f);
	x = x + tmp0;
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
	return invalid ? bitsf(0xffffffffu) : x;
}

float orc_exp_ps(float x)
{
	x = ssemin(x, 88.3762626647949f);
	x = ssemax(x, -88.3762626647949f);
	float fx = x * 1.44269504088896341f;
	fx = fx + 0.5f;
	float tmp = (float)cvtt(fx);
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
	const int n = (int)((uint32_t)(cvtt(fx) + 0x7f) << 23);
	return y * bitsf((uint32_t)n);
}

/* GammaAdj, shadertoy-util.h:186-190 */
static float gamma_adj(float c, float gamma) { return orc_exp_ps(gamma*orc_log_ps(c)); }

/* one lane of ToPixel4 (shadertoy-util.h:136-146) and ToPixel4_NoConv (148-157) */
static uint32_t to_chan(float c)
{
	int v = cvtn(255.f*c);
	if (v < 0) v = 0;
	if (v > 65535) v = 65535;
	return (v > 32767) ? 0u : (uint32_t)(v > 255 ? 255 : v);
}
static uint32_t to_chan_noconv(float c)
{
	int v = cvtn(c);
	if (v < 0) v = 0;
	if (v > 65535) v = 65535;
	return (v > 32767) ? 0u : (uint32_t)(v > 255 ? 255 : v);
}
static uint32_t to_pixel(const float c[4]) { return to_chan(c[0]) | (to_chan(c[1]) << 8) | (to_chan(c[2]) << 16) | (to_chan(c[3]) << 24); }

uint32_t orc_gamma_pixel(const float color[4], float gamma)
{
	float c[4];
	for (int i = 0; i < 4; ++i) c[i] = gamma_adj(color[i], gamma);
	return to_pixel(c);
}

/* ExpFog, shadertoy-util.h:237-241 */
static float exp_fog(float distance, float scale) { return 1.f - (expf(-scale*distance*distance*distance)); }

/* Q3_rsqrtf<2>, q3-rsqrt.h:22-42 */
static float q3_rsqrtf2(float x)
{
	const float half = 0.5f*x;
	int32_t iX; memcpy(&iX, &x, 4);
	iX = 0x5f3759df - (iX >> 1);
	memcpy(&x, &iX, 4);
	x = x*(1.5f - half*x*x);
	x = x*(1.5f - half*x*x);
	return x;
}

typedef struct { float x, y, z, w; } vec4;

static float dp4(vec4 a, vec4 b) { return (a.x*b.x + a.y*b.y) + (a.z*b.z + a.w*b.w); }  /* _mm_dp_ps(a, b, 0xff) */
static float dot3(vec4 a, vec4 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }                /* Vector3::Dot, Vector3.h:21-24 */
static void fast_norm(vec4 *v) { const float r = orc_rsqrt(dp4(*v, *v)); v->x *= r; v->y *= r; v->z *= r; v->w *= r; } /* vNorm4, shadertoy-util.h:78-92 */
static float fast_len3(vec4 v) { return sqrtf(dp4(v, v)); }                              /* vFastLen3, shadertoy-util.h:69-76 */

/* MichielPal / Desaturate, shadertoy-util.h:193-198, 244-249 */
static vec4 michiel_pal(float phase) { vec4 r = { .1f - orc_lutcosf(phase/3.f)/(19.f*0.5f), .1f, .1f + orc_lutcosf(phase/14.f)/4.f, 0.f }; return r; }
static vec4 desaturate(vec4 c, float amount)
{
	const vec4 w = { 0.0722f, 0.7152f, 0.2126f, 0.f };
	const float luma = dp4(w, c);
	vec4 r = { c.x + amount*(luma-c.x), c.y + amount*(luma-c.y), c.z + amount*(luma-c.z), c.w + amount*(luma-c.w) };
	return r;
}

static void rot2(float cosine, float sine, float *A, float *B) { const float a = cosine**A + sine**B, b = -sine**A + cosine**B; *A = a; *B = b; } /* rotY / rotZ */
static void rotX_(float angle, float *Y, float *Z) /* shadertoy-util.h:31-39 */
{
	const float cosine = orc_lutcosf(angle), sine = orc_lutsinf(angle);
	const float rY = cosine**Y + -sine**Z, rZ = sine**Y + cosine**Z;
	*Y = rY; *Z = rZ;
}
static void rotYZ_(float angle, float *A, float *B) { rot2(orc_lutcosf(angle), orc_lutsinf(angle), A, B); } /* rotY, rotZ: shadertoy-util.h:41-59 */

/* ToUV_FxMap, shadertoy-util.h:123-130 */
static void to_uv(unsigned iX, unsigned iY, float scale, float *u, float *v)
{
	float fX = (float)iX, fY = (float)iY;
	fX *= 1.f/(float)s_fxX;
	fY *= 1.f/(float)s_fxY;
	*u = (fX-0.5f)*scale*s_oneOverAspect;
	*v = (fY-0.5f)*scale;
}

static float vlerp(float a, float b, float f) { return a + f*(b-a); } /* vLerp4, shadertoy-util.h:62-67 */

/* ------------------------------------------------------------------------------------------------------------------
 * Fx_Blit_2x2 -- fx-blitter.cpp:27-75
 * ---------------------------------------------------------------------------------------------------------------- */

static uint32_t avg4(uint32_t a, uint32_t b) /* pavgb */
{
	uint32_t r = 0;
	for (int i = 0; i < 32; i += 8)
		r |= ((((a >> i) & 0xff) + ((b >> i) & 0xff) + 1) >> 1) << i;
	return r;
}

void orc_fx_blit_2x2(uint32_t *pDest, const uint32_t *pSrc)
{
	for (int iY = 0; iY < s_fxY-4; ++iY)
		for (int iX = 0; iX < s_fxX-4; ++iX)
		{
			const uint32_t r0c0 = pSrc[iY*s_fxX + iX], r0c1 = pSrc[iY*s_fxX + iX + 1];
			const uint32_t r1c0 = pSrc[(iY+1)*s_fxX + iX], r1c1 = pSrc[(iY+1)*s_fxX + iX + 1];
			const uint32_t avgH0 = avg4(r0c0, r0c1), avgH1 = avg4(r1c0, r1c1);
			uint32_t *top = pDest + (size_t)(iY*2)*s_resX + iX*2;
			top[0] = r0c0;
			top[1] = avgH0;
			top[s_resX] = avg4(r0c0, r1c0);
			top[s_resX+1] = avg4(avgH0, avgH1);
		}
}

/* ------------------------------------------------------------------------------------------------------------------
 * bilinear samplers -- bilinear.h:10-125
 * ---------------------------------------------------------------------------------------------------------------- */

static uint32_t lerp_argb(uint32_t a, uint32_t b, uint32_t f) /* ((a<<8) + (b-a)*f) >> 8 per 8-bit channel in 16-bit lanes */
{
	uint32_t r = 0;
	for (int i = 0; i < 32; i += 8)
	{
		const int ca = (int)((a >> i) & 0xff), cb = (int)((b >> i) & 0xff);
		r |= ((uint32_t)(((ca << 8) + (cb-ca)*(int)f) & 0xffff) >> 8) << i;
	}
	return r;
}

static uint32_t bsamp32(const uint32_t *tex, unsigned i00, unsigned i10, unsigned i01, unsigned i11, uint32_t fu, uint32_t fv)
{
	return lerp_argb(lerp_argb(tex[i00], tex[i10], fu), lerp_argb(tex[i01], tex[i11], fu), fv);
}

static unsigned bsamp8(const uint8_t *tex, unsigned i00, unsigned i10, unsigned i01, unsigned i11, int fu, int fv)
{
	const int S0 = tex[i00], S1 = tex[i10], S2 = tex[i01], S3 = tex[i11];
	const int S01 = ((S0 << 8) + (S1-S0)*fu) >> 8;
	const int S23 = ((S2 << 8) + (S3-S2)*fu) >> 8;
	return (unsigned)(((S01 << 8) + (S23-S01)*fv) >> 8);
}

typedef struct { unsigned i00, i10, i01, i11; uint32_t fu, fv; } texc;
static texc prep_uvs(int U, int V, unsigned mapAnd, unsigned mapShift) /* bsamp_prepUVs */
{
	unsigned U0 = (unsigned)(U >> 8), V0 = (unsigned)(V >> 8), U1 = U0+1, V1 = V0+1;
	U0 &= mapAnd; V0 = (V0 & mapAnd) << mapShift; U1 &= mapAnd; V1 = (V1 & mapAnd) << mapShift;
	texc t = { U0+V0, U1+V0, U0+V1, U1+V1, (uint32_t)(U & 0xff), (uint32_t)(V & 0xff) };
	return t;
}

/* ------------------------------------------------------------------------------------------------------------------
 * polar remap -- polar.cpp:18-198
 * ---------------------------------------------------------------------------------------------------------------- */

static void polar_maps(int32_t *pDest, int32_t *pInvDest, unsigned srcResX, unsigned srcResY) /* CalculateMaps, polar.cpp:18-59 with src == dest resolution */
{
	const float halfResX = srcResX/2.f, halfResY = srcResY/2.f;
	size_t iPixel = 0;
	const float maxDist = sqrtf(halfResX*halfResX + halfResY*halfResY);
	for (float Y = -halfResY; Y < halfResY; Y += 1.f)
		for (float X = -halfResX + KEPSILON; X < halfResX; X += 1.f)
		{
			const float distance = sqrtf(X*X + Y*Y) / maxDist;
			float theta = atan2f(Y, X);
			theta += KPI;
			theta /= KPI*2.f;
			const float U = distance*(srcResX-1.f), invU = (1.f-distance)*(srcResX-1.f), V = theta*(srcResY-1.f);
			pDest[iPixel] = (U >= srcResX-1.f) ? (int32_t)(((srcResX-2)<<8) | 0xff) : ftofp24(U);
			pInvDest[iPixel] = (invU >= srcResX-1.f) ? (int32_t)(((srcResX-2)<<8) | 0xff) : ftofp24(invU);
			pInvDest[iPixel+1] = pDest[iPixel+1] = (V >= srcResY-1.f) ? (int32_t)(((srcResY-2)<<8) | 0xff) : ftofp24(V);
			iPixel += 2;
		}
}

void orc_polar_maps(int32_t *pDest, int32_t *pInvDest) { polar_maps(pDest, pInvDest, (unsigned)s_resX, (unsigned)s_resY); }      /* polar.cpp:68 */
void orc_polar_maps_2x2(int32_t *pDest, int32_t *pInvDest) { polar_maps(pDest, pInvDest, (unsigned)s_fxX, (unsigned)s_fxY); }  /* polar.cpp:69 */

static void polar_blit(uint32_t *pDest, const uint32_t *pSrc, const int32_t *pMap, int alpha, unsigned resX, unsigned resY)
{
	const size_t n = (size_t)resX*resY;
	for (size_t i = 0; i < n; ++i)
	{
		const int U = pMap[i*2], V = pMap[i*2+1];
		const unsigned U0 = (unsigned)(U >> 8), V0 = (unsigned)(V >> 8)*resX;
		const uint32_t s = bsamp32(pSrc, U0+V0, U0+1+V0, U0+V0+resX, U0+1+V0+resX, U & 0xff, V & 0xff);
		pDest[i] = alpha ? lerp_argb(pDest[i], s, s >> 24) : s; /* polar.cpp:169-174 */
	}
}

void orc_polar_blit(uint32_t *pDest, const uint32_t *pSrc, const int32_t *pMap, int alpha) /* Polar_Blit / Polar_BlitA, polar.cpp:135-198 */
{
	polar_blit(pDest, pSrc, pMap, alpha, (unsigned)s_resX, (unsigned)s_resY);
}

/* Polar_Blit_2x2, polar.cpp:200-218: the tile walk restated as what it writes inside the FX map (the tiles' overrun past
 * the last row and the right edge is the reference's defect; the oracle build clamps it: P2/P2b) */
void orc_polar_blit_2x2(uint32_t *pDest, const uint32_t *pSrc, const int32_t *pMap)
{
	polar_blit(pDest, pSrc, pMap, 0, (unsigned)s_fxX, (unsigned)s_fxY);
}

/* ------------------------------------------------------------------------------------------------------------------
 * old box blur -- deprecated/boxblur.cpp:11-231
 * ---------------------------------------------------------------------------------------------------------------- */

static unsigned weight_to_div16(unsigned weight) { return (((65536u*256u)/weight) >> 4) & 0xffffu; } /* WeightToDiv into a 16-bit lane */
static unsigned sat16(unsigned v) { return v > 65535u ? 65535u : v; }
static unsigned subs16(unsigned a, unsigned b) { return a > b ? a - b : 0u; }

static void old_blur_line(uint8_t *dst, const uint8_t *src, size_t step, unsigned len, float strength)
{
	unsigned kernelSpan = f2u(strength*255.f);
	kernelSpan = (unsigned)((int)kernelSpan < 1 ? 1 : ((int)kernelSpan > 255 ? 255 : (int)kernelSpan));
	const int subEdges = (kernelSpan & 1) == 0;
	const unsigned edgeSpan = kernelSpan >> 1;
	const unsigned remainderShift = 1 + ((!subEdges)*7);
	const unsigned kernelMedian = edgeSpan + !subEdges;
	const unsigned startWeight = (kernelMedian << 4) + ((unsigned)subEdges << 3);
	const unsigned fullPassLen = len - (kernelMedian+edgeSpan);
	const unsigned fullDiv = weight_to_div16(kernelSpan << 4);

	for (int c = 0; c < 4; ++c) /* the four 16-bit lanes are independent */
	{
		unsigned acc = 0, addRem = 0, subRem = 0;
		size_t addPos = c, subPos = c, destPos = c;
#define ADD(px) do { acc = sat16(acc + addRem); addRem = (px) >> remainderShift; acc = sat16(acc + ((px) - addRem)); } while (0)
#define SUB(px) do { acc = subs16(acc, subRem); subRem = (px) >> remainderShift; acc = subs16(acc, (px) - subRem); } while (0)
#define DIV(d) do { const unsigned v = (acc*(d)) >> 16; dst[destPos] = (uint8_t)((v > 32767u) ? 0u : (v > 255u ? 255u : v)); destPos += step; } while (0)
		for (unsigned i = 0; i < edgeSpan; ++i) { const unsigned px = src[addPos]; addPos += step; ADD(px); }
		for (unsigned i = 0; i < kernelMedian; ++i) { const unsigned px = src[addPos]; addPos += step; ADD(px); DIV(weight_to_div16(startWeight + 16*i)); }
		for (unsigned i = 0; i < fullPassLen; ++i)
		{
			const unsigned px = src[addPos]; addPos += step; ADD(px);
			const unsigned spx = src[subPos]; subPos += step; SUB(spx); /* in place: a pixel this loop already wrote */
			DIV(fullDiv);
		}
		if (subEdges) acc = sat16(acc + addRem);
		for (unsigned i = edgeSpan; i > 0; --i) { const unsigned spx = src[subPos]; subPos += step; SUB(spx); DIV(weight_to_div16(startWeight + 16*(i-1))); }
#undef ADD
#undef SUB
#undef DIV
	}
}

void orc_old_blur_h(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength)
{
	for (unsigned y = 0; y < yRes; ++y)
		old_blur_line((uint8_t *)(pDest + (size_t)y*xRes), (const uint8_t *)(pSrc + (size_t)y*xRes), 4, xRes, strength);
}
void orc_old_blur_v(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength)
{
	for (unsigned x = 0; x < xRes; ++x)
		old_blur_line((uint8_t *)(pDest + x), (const uint8_t *)(pSrc + x), (size_t)xRes*4, yRes, strength);
}
void orc_old_blur(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength) /* BoxBlur32 */
{
	orc_old_blur_h(pDest, pSrc, xRes, yRes, strength);
	orc_old_blur_v(pDest, pDest, xRes, yRes, strength);
}
float orc_box_blur_scale(float strength) { if (strength != 0.f) strength = clampf_(1.f, 100.f, strength)*0.01f; return strength; }

/* ------------------------------------------------------------------------------------------------------------------
 * new box blur -- boxblur.cpp:87-318 (the reference's own transposed buffer layout is kept so that the 2-pixel
 * over-read past each line lands on the same data; pixels past the end of a buffer read as 0)
 * ---------------------------------------------------------------------------------------------------------------- */

static unsigned new_div_pack(int32_t iSum, uint32_t scale)
{
	const uint64_t q = ((uint64_t)(uint32_t)iSum*scale) >> 22;
	const int32_t v = (int32_t)((uint32_t)q + (uint32_t)(q >> 32));
	if (v < 0) return 0;
	const unsigned w = v > 65535 ? 65535u : (unsigned)v;
	return (w > 32767u) ? 0u : (w > 255u ? 255u : w);
}
static uint32_t scale_fp22(float v) { const float s = v*(float)(1<<22); return (s >= -9223372036854775808.f && s < 9223372036854775808.f) ? (uint32_t)(uint64_t)(int64_t)s : 0u; }

static void horz_blur32(uint32_t *pDest, uint32_t *pScratch, const uint32_t *pSrc, size_t srcElems, unsigned xRes, unsigned yRes,
	unsigned writeStrideCol, unsigned writeStrideRow, float strength, float gain, unsigned numPasses)
{
	strength *= 0.01f;
	const float radius = stdminf(500.f, strength*(float)((xRes-2)/2));
	const unsigned iSpan = f2u(radius);
	const float scale = 1.f/((2.f-gain)*radius + 1.f);
	const float alpha = radius-(float)iSpan;
	const uint32_t iScale = scale_fp22(scale);
	const int32_t iAlpha = cvtt(65536.f*alpha);
	const float halfScale = scale*0.5f, dScale = halfScale/(float)iSpan;
	const size_t elems = (size_t)xRes*yRes;
	const uint32_t *pRead = pSrc;
	size_t readElems = srcElems;
	if (0 == (numPasses & 1)) { uint32_t *t = pDest; pDest = pScratch; pScratch = t; }

	for (unsigned iPass = 0; iPass < numPasses; ++iPass)
	{
		unsigned colStride = xRes, rowStride = 1;
		if (iPass == numPasses-1) { colStride = writeStrideCol; rowStride = writeStrideRow; }
		for (unsigned iY = 0; iY < yRes; ++iY)
			for (int c = 0; c < 32; c += 8)
			{
#define PX(i) ((size_t)iY*xRes + (i) < readElems ? (int32_t)((pRead[(size_t)iY*xRes + (i)] >> c) & 0xff) : 0)
#define LERP(A, B) ((A) + ((((B)-(A))*iAlpha) >> 16))
#define OUT(s) do { uint32_t *p = pDest + writeIdx; *p = (*p & ~(0xffu << c)) | (new_div_pack(iSum, (s)) << c); writeIdx += rowStride; } while (0)
				size_t writeIdx = (size_t)iY*colStride;
				int32_t iSum = 0;
				unsigned tail = 0, head = 0;
				for (unsigned i = 0; i < iSpan; ++i) { iSum += PX(head); ++head; }
				iSum += (PX(head)*iAlpha) >> 16;
				int32_t headA = PX(head+1), headB, tailA, tailB;
				for (unsigned i = 0; i < iSpan; ++i)
				{
					OUT(scale_fp22(halfScale + (float)i*dScale));
					headB = PX(head+2); iSum += LERP(headA, headB); headA = headB; ++head;
				}
				tailA = PX(tail);
				for (unsigned i = 0; i < xRes - iSpan*2; ++i)
				{
					OUT(iScale);
					headB = PX(head+2); iSum += LERP(headA, headB); headA = headB; ++head;
					tailB = PX(tail+1); iSum -= LERP(tailA, tailB); tailA = tailB; ++tail;
				}
				for (unsigned i = iSpan; i > 0; --i)
				{
					OUT(scale_fp22(halfScale + (float)(i-1)*dScale));
					tailB = PX(tail+1); iSum -= LERP(tailA, tailB); tailA = tailB; ++tail;
				}
#undef PX
#undef LERP
#undef OUT
			}
		pRead = pDest;
		readElems = elems;
		{ uint32_t *t = pDest; pDest = pScratch; pScratch = t; }
	}
}

static void transpose32(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes) /* Transpose32, boxblur.cpp:215-268 */
{
	for (unsigned y = 0; y < yRes; ++y)
		for (unsigned x = 0; x < xRes; ++x)
			pDest[(size_t)x*yRes + y] = pSrc[(size_t)y*xRes + x];
}

/* kind: 0 = BoxBlur_Horz32, 1 = BoxBlur_Vert32, 2 = BoxBlur_32 (boxblur.cpp:270-318); srcElems = readable elements of pSrc */
void orc_new_blur(int kind, uint32_t *pDest, const uint32_t *pSrc, size_t srcElems, unsigned xRes, unsigned yRes, float strength, float gain, unsigned numPasses)
{
	const size_t elems = (size_t)xRes*yRes;
	uint32_t *s0 = (uint32_t *)calloc(elems, 4), *s1 = (uint32_t *)calloc(elems, 4);
	if (kind == 0)
		horz_blur32(pDest, s0, pSrc, srcElems, xRes, yRes, xRes, 1, strength, gain, numPasses);
	else if (kind == 1)
	{
		transpose32(s1, pSrc, xRes, yRes);
		horz_blur32(pDest, s0, s1, elems, yRes, xRes, 1, xRes, strength, gain, numPasses);
	}
	else
	{
		horz_blur32(s1, s0, pSrc, srcElems, xRes, yRes, 1, yRes, strength, gain, numPasses);
		horz_blur32(pDest, s0, s1, elems, yRes, xRes, 1, xRes, strength, gain, numPasses);
	}
	free(s0); free(s1);
}

/* ------------------------------------------------------------------------------------------------------------------
 * blend ops / blits -- util.cpp:83-812
 * ---------------------------------------------------------------------------------------------------------------- */

static unsigned soft_light(unsigned A, unsigned B) /* SoftLightBlend, util.cpp:227-240 */
{
	const int dA = (int)(A/2)+64;
	if (B < 128) return (unsigned)((2*dA*(int)B)/256);
	return (unsigned)(255 - (2*(255-dA)*(255-(int)B))/256);
}
static unsigned overlay_chan(unsigned bottom, unsigned top) { return bottom < 128 ? (2*bottom*top/255) : (255 - 2*(255-bottom)*(255-top)/255); }
static uint32_t adds4(uint32_t a, uint32_t b) { uint32_t r = 0; for (int i = 0; i < 32; i += 8) { unsigned v = ((a>>i)&0xff) + ((b>>i)&0xff); r |= (v > 255 ? 255u : v) << i; } return r; }
static uint32_t subs4(uint32_t a, uint32_t b) { uint32_t r = 0; for (int i = 0; i < 32; i += 8) { int v = (int)((a>>i)&0xff) - (int)((b>>i)&0xff); r |= (uint32_t)(v < 0 ? 0 : v) << i; } return r; }

/* op ids = ckd_blend_op (include/ckd.h) */
void orc_blend(int op, uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels, float fParam, unsigned uParam)
{
	const unsigned iA = f2u(saturatef_(fParam)*255.f); /* SoftLight32AA, util.cpp:311-312 */
	for (unsigned i = 0; i < numPixels; ++i)
	{
		const uint32_t d = pDest[i], s = (op == 14) ? 0u : pSrc[i];
		const unsigned A2 = d>>24, R2 = (d>>16)&0xff, G2 = (d>>8)&0xff, B2 = d&0xff;
		const unsigned A1 = s>>24, R1 = (s>>16)&0xff, G1 = (s>>8)&0xff, B1 = s&0xff;
		unsigned R, G, B, a;
		uint32_t out = d;
		switch (op)
		{
		case 0: out = lerp_argb(d, s, uParam & 0xff); break;                                  /* Mix32, util.cpp:83-96 */
		case 1: a = 0xff - A1;                                                                 /* MixOver32, util.cpp:148-177 */
			R = ((R1*(0xff-a))>>8) + ((R2*a)>>8); G = ((G1*(0xff-a))>>8) + ((G2*a)>>8); B = ((B1*(0xff-a))>>8) + ((B2*a)>>8);
			if (R>255) R=255;
			if (G>255) G=255;
			if (B>255) B=255;
			out = (R<<16)|(G<<8)|B; break;
		case 2: out = adds4(d, s); break;                                                      /* Add32 */
		case 3: out = subs4(d, s); break;                                                      /* Sub32 */
		case 4: R = R1+R2-((2*R1*R2)>>8); G = G1+G2-((2*G1*G2)>>8); B = B1+B2-((2*B1*B2)>>8);  /* Excl32 */
			out = (A2<<24)|(R<<16)|(G<<8)|B; break;
		case 5: out = (A2<<24)|(soft_light(R1,R2)<<16)|(soft_light(G1,G2)<<8)|soft_light(B1,B2); break; /* SoftLight32 */
		case 6: case 7:                                                                        /* SoftLight32A / AA: unsigned lerp, bits spill */
			a = (op == 6) ? A1 : iA;
			R = soft_light(R1,R2); G = soft_light(G1,G2); B = soft_light(B1,B2);
			R = R2+(((R-R2)*a)>>8); G = G2+(((G-G2)*a)>>8); B = B2+(((B-B2)*a)>>8);
			out = ((op == 7) ? (a<<24) : 0u)|(R<<16)|(G<<8)|B; break;
		case 8: out = (overlay_chan(R2,R1)<<16)|(overlay_chan(G2,G1)<<8)|overlay_chan(B2,B1); break; /* Overlay32 */
		case 9: R = overlay_chan(R2,R1); G = overlay_chan(G2,G1); B = overlay_chan(B2,B1);     /* Overlay32A */
			R = R2+(((R-R2)*A1)>>8); G = G2+(((G-G2)*A1)>>8); B = B2+(((B-B2)*A1)>>8);
			out = (R<<16)|(G<<8)|B; break;
		case 10: R = (R2+(R1<R2?R1:R2))>>1; G = (G2+(G1<G2?G1:G2))>>1; B = (B2+(B1<B2?B1:B2))>>1; /* Darken32_50 */
			out = (A2<<24)|(R<<16)|(G<<8)|B; break;
		case 11: out = (((A1*A2)>>8)<<24)|(((R1*R2)>>8)<<16)|(((G1*G2)>>8)<<8)|((B1*B2)>>8); break; /* MulSrc32 */
		case 12: out = (((A1*A2)>>8)<<24)|(((A1*R2)>>8)<<16)|(((A1*G2)>>8)<<8)|((A1*B2)>>8); break; /* MulSrc32A */
		case 13: out = lerp_argb(d, s, A1); break;                                             /* MixSrc32 */
		case 14: out = lerp_argb(d, uParam & 0xffffff, uParam >> 24); break;                   /* Fade32 */
		default: break;
		}
		pDest[i] = out;
	}
}

/* op ids = ckd_blit_op; MixSrc32S = op 0 with srcStride */
void orc_blit(int op, uint32_t *pDest, const uint32_t *pSrc, unsigned destStride, unsigned srcStride, unsigned width, unsigned yRes, float alpha)
{
	const uint32_t fa = 0x01010101u*f2u(alpha*255.f);
	for (unsigned y = 0; y < yRes; ++y)
		for (unsigned x = 0; x < width; ++x)
		{
			uint32_t *pd = pDest + (size_t)y*destStride + x;
			const uint32_t d = *pd, s = pSrc[(size_t)y*srcStride + x];
			uint32_t out = 0;
			if (op == 0) out = lerp_argb(d, s, s >> 24);
			else if (op == 2) out = adds4(d, s);
			else for (int c = 0; c < 32; c += 8)
			{
				const unsigned f = (fa >> c) & 0xff, sc = (s >> c) & 0xff, dc = (d >> c) & 0xff;
				unsigned r;
				if (op == 1) { const unsigned a = ((s >> 24)*f) >> 8; r = (((dc << 8) + a*(sc - dc)) & 0xffff) >> 8; if (r > 255) r = 255; }
				else { r = dc + ((sc*f) >> 8); if (r > 255) r = 255; }
				out |= r << c;
			}
			*pd = out;
		}
}

/* TapeWarp32, util.cpp:552-603 */
void orc_tape_warp(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float speed)
{
	for (int iY = 0; iY < (int)yRes; ++iY)
		for (unsigned iX = 0; iX < xRes; ++iX)
		{
			const float dX = orc_lutsinf(iY*speed)*strength*1.f;
			const float dY = orc_lutcosf(iX*speed)*strength*1.f;
			float tX = iX + dX, tY = iY + dY;
			if (tX < 0.f) tX = 0.f; else if (tX >= xRes-1.f) tX = xRes - 2.f;
			if (tY < 0.f) tY = 0.f; else if (tY >= yRes-1.f) tY = yRes - 2.f;
			const int U = ftofp24(tX), V = ftofp24(tY);
			const unsigned U0 = (unsigned)(U >> 8), V0 = (unsigned)(V >> 8)*(unsigned)s_resX;
			pDest[(size_t)iY*xRes + iX] = bsamp32(pSrc, U0+V0, U0+1+V0, U0+V0+s_resX, U0+1+V0+s_resX, U & 0xff, V & 0xff);
		}
}

/* ------------------------------------------------------------------------------------------------------------------
 * raymarchers -- shadertoy.cpp (FX map only; the caller applies orc_fx_blit_2x2 and the post chain)
 * ---------------------------------------------------------------------------------------------------------------- */

static float fPlasma(vec4 p, float time) /* shadertoy.cpp:201-209 */
{
	const float sine = 0.2f*orc_lutsinf(p.x-p.y);
	const float fX = sine + orc_lutcosf(p.x*0.33f), fY = sine + orc_lutcosf(p.y*0.43f), fZ = sine + orc_lutcosf((5.f*time+p.z)*0.53f);
	return sqrtf(fX*fX + fY*fY + fZ*fZ)-0.8f;
}

void orc_plasma_map(uint32_t *pDest, float time, float speed, float hue, float gamma, float desaturation) /* RenderPlasmaMap, shadertoy.cpp:211-274 */
{
	const vec4 colMulA = desaturate(michiel_pal(hue), desaturation);
	const vec4 colMulB = desaturate(colMulA, 0.8f);
	time = time*speed;
	const float angle = time*0.314f*0.5f, dirCos = orc_lutcosf(angle), dirSin = orc_lutsinf(angle);
	for (int iY = 0; iY < s_fxY; ++iY)
		for (int iX = 0; iX < s_fxX; ++iX)
		{
			float u, v; to_uv(iX, iY, 4.f, &u, &v);
			const vec4 dir = { dirCos*u*s_aspect - dirSin*0.75f, v, dirSin*u + dirCos*0.75f, 0.f };
			float total = 0.f, march = 0.f;
			vec4 hit = { 0.f, 0.f, 0.f, 0.f };
			for (int iStep = 0; iStep < 24; ++iStep)
			{
				march = fPlasma(hit, time);
				total += march*(0.5f*KGOLDENRATIO);
				hit.x = dir.x*total; hit.y = dir.y*total; hit.z = dir.z*total;
			}
			const vec4 half = { hit.x*0.5f, hit.y*0.5f, hit.z*0.5f, 0.f };
			const float second = fPlasma(half, time), mul = 8.f - dir.x*0.5f;
			const float color[4] = { (colMulA.x*march + colMulB.x*second)*mul, (colMulA.y*march + colMulB.y*second)*mul, (colMulA.z*march + colMulB.z*second)*mul, 0.f };
			pDest[iY*s_fxX + iX] = orc_gamma_pixel(color, gamma);
		}
}

static vec4 s_nautilusGlobal;
static float fNautilus(vec4 p, float time) /* shadertoy.cpp:289-298 */
{
	const float cosX = orc_lutcosf(orc_lutcosf(p.x + s_nautilusGlobal.x)*p.x - orc_lutcosf(p.y + s_nautilusGlobal.y)*p.y);
	const float cosY = orc_lutcosf(p.z*0.33f*p.x - s_nautilusGlobal.z*p.y);
	const float cosZ = orc_lutcosf(p.x + p.y + p.z*0.8f + time);
	return (cosX*cosX + cosY*cosY + cosZ*cosZ)*0.5f - .7f;
}
static vec4 offs(vec4 p, float dx, float dy, float dz) { vec4 r = { p.x+dx, p.y+dy, p.z+dz, 0.f }; return r; }

void orc_nautilus_map(uint32_t *pDest, float time, float roll, float hue, float speed, float desaturation) /* shadertoy.cpp:300-393 */
{
	time = time*speed;
	s_nautilusGlobal.x = time*0.125f; s_nautilusGlobal.y = time/9.f; s_nautilusGlobal.z = orc_lutcosf(time*0.1428f);
	const vec4 colorization = { .1f-orc_lutcosf(hue/3.f)/19.f, .1f, .1f+orc_lutcosf(hue/14.f)/8.f, 0.f };
	const vec4 diffColor = desaturate(colorization, desaturation);
	const float cosHitOffs = orc_lutcosf(time*0.314f*0.5f), funkCos = orc_lutcosf(time*KGOLDENRATIO*0.1f);
	for (int iY = 0; iY < s_fxY; ++iY)
		for (int iX = 0; iX < s_fxX; ++iX)
		{
			float u, v; to_uv(iX, iY, 2.f, &u, &v);
			vec4 dir = { u*s_aspect, v, 1.f, 0.f };
			rotYZ_(roll*time, &dir.x, &dir.y);
			fast_norm(&dir);
			vec4 hit = { 0.f, 0.f, 0.f, 0.f };
			float total = 0.01f, march = 1.f;
			for (int iStep = 0; march > 0.01f && iStep < 48; ++iStep)
			{
				hit.x = dir.x*total; hit.y = dir.y*total; hit.z = dir.z*total;
				march = fNautilus(hit, time);
				total += march*0.628f;
			}
			const float nOffs = 0.15f;
			vec4 normal = { march-fNautilus(offs(hit, nOffs, 0, 0), time), march-fNautilus(offs(hit, 0, nOffs, 0), time), march-fNautilus(offs(hit, 0, 0, nOffs), time), 0.f };
			fast_norm(&normal);
			float diffuse = normal.z*0.1f;
			const float specular = powf(stdmaxf(0.f, dot3(normal, dir)), 16.f);
			const vec4 hitOffs = offs(hit, cosHitOffs, cosHitOffs, cosHitOffs);
			vec4 funk = { march-fNautilus(offs(hitOffs, nOffs, 0, 0), time), march-fNautilus(offs(hitOffs, 0, nOffs, 0), time), march-fNautilus(offs(hitOffs, 0, 0, nOffs), time), 0.f };
			fast_norm(&funk);
			const float yMod = fracf_(hit.y*0.3f + funk.x*0.628f + funk.y*funkCos);
			diffuse *= yMod*yMod*yMod;
			const float s = 1.56f*total + specular, add = specular*KGOLDENRATIO*0.2f;
			const float color[4] = { (diffuse + diffColor.x*s) + add, (diffuse + diffColor.y*s) + add, (diffuse + diffColor.z*s) + add, 0.f };
			pDest[iY*s_fxX + iX] = orc_gamma_pixel(color, 1.44f);
		}
}

static vec4 s_spikeGlobal;
static float fSpikey(vec4 p, float scaleBase) /* fSpikey1 / fSpikey2, shadertoy.cpp:418-430 */
{
	const float scale = scaleBase*0.1f;
	const float radius = 1.35f + scale*orc_lutcosf(s_spikeGlobal.y*p.y - s_spikeGlobal.x) + scale*orc_lutcosf(s_spikeGlobal.z*p.x + s_spikeGlobal.x);
	return fast_len3(p) - radius;
}

/* variant: 0 = close (shadertoy.cpp:432-523), 1 = distant (525-598), 2 = specular only (600-659).
 * p[]: speed, roll, specPow, desaturation, hue, gamma, xOffs, yOffs, zOffs, zOffsScale, normalGrain, scale, aspectMul, warmup */
void orc_spikey_map(int variant, uint32_t *pDest, float time, const float *p)
{
	const float speed = p[0], roll = p[1], specPow = p[2], desaturation = p[3], hue = p[4], gamma = p[5];
	const float xOffs = p[6], yOffs = p[7], zOffs = p[8], zOffsScale = p[9], normalGrain = p[10], scale = p[11], warmup = p[13];
	const int aspectMul = p[12] != 0.f;
	const vec4 diffColor = desaturate(michiel_pal(hue), desaturation);
	const float dc[4] = { diffColor.x, diffColor.y, diffColor.z, diffColor.w };
	float zOffsFinal = 0.f;
	if (variant == 0)
	{
		/* easeInOutElasticf, synth-math-easings.h:190-203 */
		const float c5 = (2.f*KPI)/4.5f, x = zOffs;
		const float e = (0.f == x) ? 0.f : (1.f == x) ? 1.f : (x < 0.5f)
			? -(powf(2.f, 20.f*x - 10.f) * sinf((20.f*x - 11.125f) * c5))*0.5f
			: (powf(2.f, -20.f*x + 10.f) * sinf((20.f*x - 11.125f) * c5))*0.5f + 1.f;
		zOffsFinal = e*zOffsScale;
		s_spikeGlobal.x = speed*time; s_spikeGlobal.y = 16.f*scale; s_spikeGlobal.z = aspectMul ? s_aspect*22.f*scale : 22.f*scale;
	}
	else if (variant == 1) { s_spikeGlobal.x = speed*time; s_spikeGlobal.y = 16.f; s_spikeGlobal.z = 16.f; }
	else { s_spikeGlobal.x = speed*time; s_spikeGlobal.y = 8.f; s_spikeGlobal.z = 16.f; }

	for (int iY = 0; iY < s_fxY; ++iY)
		for (int iX = 0; iX < s_fxX; ++iX)
		{
			float u, v; to_uv(iX, iY, variant == 2 ? KGOLDENRATIO : 2.f, &u, &v);
			vec4 origin, dir;
			if (variant == 0) { vec4 o = { 0.2f, 0.f, -2.23f, 0.f }, d = { (u+xOffs)*s_aspect, v + yOffs, 1.f + zOffsFinal, 0.f }; origin = o; dir = d; }
			else if (variant == 1) { vec4 o = { 0.f, 0.f, -2.614f + zOffs, 0.f }, d = { u + xOffs, v + yOffs, 1.f, 0.f }; origin = o; dir = d; }
			else { vec4 o = { 0.f, 0.f, -3.314f, 0.f }, d = { u*s_aspect, v, 1.f, 0.f }; origin = o; dir = d; }
			rotYZ_(roll, &dir.x, &dir.y);
			fast_norm(&dir);
			vec4 hit = { 0.f, 0.f, 0.f, 0.f };
			float march = 1.f, total = 0.f;
			const float base = (variant == 0) ? KGOLDENANGLE : KGOLDENRATIO;
			if (variant == 0)
				for (int iStep = 0; march > 0.0001f && iStep < 32; ++iStep)
				{
					hit.x = origin.x + dir.x*total; hit.y = origin.y + dir.y*total; hit.z = origin.z + dir.z*total;
					march = fSpikey(hit, base);
					total += march*(0.05f*KPI);
				}
			else if (variant == 1)
				for (int iStep = 0; march > 0.001f && iStep < 48; ++iStep)
				{
					hit.x = origin.x + dir.x*total; hit.y = origin.y + dir.y*total; hit.z = origin.z + dir.z*total;
					march = fSpikey(hit, base);
					march *= 0.314f;
					total += march;
				}
			else
				for (int iStep = 0; iStep < 36; ++iStep)
				{
					hit.x = origin.x + dir.x*total; hit.y = origin.y + dir.y*total; hit.z = origin.z + dir.z*total;
					march = fSpikey(hit, base);
					total += march*0.075f*KGOLDENRATIO;
				}
			const float nOffs = (variant == 0) ? normalGrain : (variant == 1) ? KPI*0.02f : 0.01f;
			vec4 normal = { march-fSpikey(offs(hit, nOffs, 0, 0), base), march-fSpikey(offs(hit, 0, nOffs, 0), base), march-fSpikey(offs(hit, 0, 0, nOffs), base), 0.f };
			fast_norm(&normal);
			const float distance = hit.z-origin.z;
			float color[4];
			if (variant == 2)
			{
				const float fakeSpecular = warmup*powf(stdmaxf(0.f, dot3(normal, dir)), specPow);
				const float fogged = vlerp(fakeSpecular, 0.f, exp_fog(distance, 0.0133f));
				pDest[iY*s_fxX + iX] = to_chan(fogged)*0x01010101u;
				continue;
			}
			float diffuse, specular, fog;
			if (variant == 0)
			{
				diffuse = normal.z; /* rim (shadertoy.cpp:504-511) multiplies by max(1, min(0, rim)) == 1 */
				specular = powf(stdmaxf(0.f, dot3(normal, dir)), specPow);
				fog = exp_fog(distance, KGOLDENRATIO*0.1f);
			}
			else
			{
				diffuse = stdmaxf(0.f, normal.z*0.8f + normal.y*0.2f);
				specular = powf(dot3(normal, dir), specPow);
				fog = exp_fog(distance, 0.133f);
			}
			for (int i = 0; i < 4; ++i) color[i] = vlerp((dc[i] + specular)*diffuse, 1.f, fog);
			pDest[iY*s_fxX + iX] = orc_gamma_pixel(color, gamma);
		}
}

static float fSinMap(vec4 point) /* shadertoy.cpp:879-899 */
{
	const float pZ = point.z, zMod = pZ*0.314f;
	const float pathCos = orc_lutcosf(zMod), pathCos2 = orc_lutcosf(zMod+(K2PI/4.f))*KGOLDENRATIO;
	const float pX = point.x-(pathCos2*2.f - pathCos*1.5f), pY = point.y-(pathCos*3.14f + pathCos2);
	const float aX = pX*0.315f*1.25f + orc_lutsinf(pZ*(0.814f*1.25f));
	const float aY = pY*0.315f*1.25f + orc_lutsinf(pX*(0.814f*1.25f));
	const float aZ = pZ*0.315f*1.25f + orc_lutsinf(pY*(0.814f*1.25f));
	const float cosX = orc_lutcosf(aX), cosY = orc_lutcosf(aY), cosZ = orc_lutcosf(aZ);
	return (sqrtf(cosX*cosX + cosY*cosY + cosZ*cosZ) - 1.025f)*1.33f;
}

void orc_sinuses_map(uint32_t *pDest, float time, float specular, float roll, float speed, float offsX, float gamma, float hue, float desaturation) /* shadertoy.cpp:901-982 */
{
	const float specPow = 1.f + specular;
	const vec4 diffColor = desaturate(michiel_pal(hue), desaturation);
	const float dc[4] = { diffColor.x, diffColor.y, diffColor.z, diffColor.w };
	const float pathTime = time*speed, timeMod = pathTime*0.314f, sine = orc_lutsinf(timeMod), cosine = orc_lutcosf(timeMod); /* fSinPath */
	const vec4 origin = { sine*2.f*KGOLDENRATIO - cosine*1.5f, cosine*3.14f + sine*KGOLDENRATIO, pathTime, 0.f };
	for (int iY = 0; iY < s_fxY; ++iY)
		for (int iX = 0; iX < s_fxX; ++iX)
		{
			float u, v; to_uv(iX, iY, 2.f, &u, &v);
			vec4 dir = { (u+offsX)*s_aspect, v, 0.314f, 0.f };
			rotYZ_(roll, &dir.x, &dir.y);
			fast_norm(&dir);
			vec4 hit = { 0.f, 0.f, 0.f, 0.f };
			float march = 1.f, total = 0.f;
			for (int iStep = 0; march > 0.01f && iStep < 32; ++iStep)
			{
				hit.x = origin.x + dir.x*total; hit.y = origin.y + dir.y*total; hit.z = origin.z + dir.z*total;
				march = fSinMap(hit);
				total += march*0.814f;
			}
			const float nOffs = 0.2f;
			vec4 normal = { march-fSinMap(offs(hit, nOffs, 0, 0)), march-fSinMap(offs(hit, 0, nOffs, 0)), march-fSinMap(offs(hit, 0, 0, nOffs)), 0.f };
			fast_norm(&normal);
			float diffuse = normal.z*0.7f + 0.3f*normal.y;
			diffuse = 0.2f + 0.8f*diffuse;
			const float fakeSpecular = powf(dot3(normal, dir), specPow);
			const float fog = exp_fog(hit.z-origin.z, 0.03f);
			float color[4];
			for (int i = 0; i < 4; ++i) color[i] = vlerp((dc[i] + fakeSpecular)*diffuse, 1.f, fog);
			pDest[iY*s_fxX + iX] = orc_gamma_pixel(color, gamma);
		}
}

static float fLaura(vec4 p) { return orc_lutcosf(p.x)+orc_lutcosf(p.y)+orc_lutcosf(p.z) + 1.f; } /* shadertoy.cpp:998-1001 */

void orc_laura_map(uint32_t *pDest, float time, float speed, float yaw, float pitch, float roll, float hue, float saturate) /* shadertoy.cpp:1017-1103 */
{
	const vec4 diffColor = desaturate(michiel_pal(hue), saturate);
	const float dc[4] = { diffColor.x, diffColor.y, diffColor.z, diffColor.w };
	const vec4 origin = { 0.f, 0.f, speed*time, 0.f };
	for (int iY = 0; iY < s_fxY; ++iY)
		for (int iX = 0; iX < s_fxX; ++iX)
		{
			float u, v; to_uv(iX, iY, 2.f, &u, &v);
			vec4 dir = { u*s_aspect, v, KPI, 0.f };
			rotYZ_(yaw, &dir.x, &dir.z);
			rotX_(pitch, &dir.y, &dir.z);
			rotYZ_(roll*time, &dir.x, &dir.y);
			fast_norm(&dir);
			vec4 hit = { 0.f, 0.f, 0.f, 0.f };
			float march = 0.f, total = 0.f;
			for (int iStep = 0; iStep < 32; ++iStep)
			{
				hit.x = origin.x + dir.x*total; hit.y = origin.y + dir.y*total; hit.z = origin.z + dir.z*total;
				march = fLaura(hit);
				total += march*0.5f;
			}
			const float nOffs = 0.1628f; /* LauraNormal, shadertoy.cpp:1003-1015 */
			vec4 normal = { fLaura(offs(hit, nOffs, 0, 0))-march, fLaura(offs(hit, 0, nOffs, 0))-march, fLaura(offs(hit, 0, 0, nOffs))-march, 0.f };
			fast_norm(&normal);
			const vec4 lightPos = { origin.x-dir.x, origin.y-dir.y, origin.z-dir.z, 0.f };
			vec4 lightDir = { lightPos.x-hit.x, lightPos.y-hit.y, lightPos.z-hit.z, 0.f };
			fast_norm(&lightDir);
			float diffuse = stdmaxf(0.3f, dot3(normal, lightDir));
			const float distance = hit.z-origin.z;
			/* Shadertoy::Specular, shadertoy-util.h:252-280 */
			vec4 V = { origin.x-hit.x, origin.y-hit.y, origin.z-hit.z, origin.w-hit.w };
			fast_norm(&V);
			vec4 H = { lightDir.x+V.x, lightDir.y+V.y, lightDir.z+V.z, lightDir.w+V.w };
			fast_norm(&H);
			const float cosAng = dp4(normal, H);
			const float specular = (0 == (fbits(cosAng) >> 31)) ? powf(cosAng, 4.f) : 0.f;
			float rim = diffuse*diffuse;
			rim = (rim*rim-0.13f)*32.f;
			rim = stdmaxf(1.f, stdminf(0.f, rim));
			diffuse *= rim;
			const float fogColor = q3_rsqrtf2(specular+diffuse), fog = exp_fog(distance, 0.001f);
			float color[4];
			for (int i = 0; i < 4; ++i) color[i] = vlerp(dc[i]*(diffuse+specular), fogColor, fog);
			pDest[iY*s_fxX + iX] = orc_gamma_pixel(color, 1.44f);
		}
}

/* RenderTunnelMap_2x2, shadertoy.cpp:746-838.  p[]: boxy, flowerScale, flowerFreq, flowerPhase, speed, roll, pitch, radius, uMul, vMul, fog1, fog2 */
void orc_tunnel_map(uint32_t *pDest, uint32_t *pGlowDest, const uint32_t *tex, const uint32_t *texGlow, float time, const float *p)
{
	const float boxy = p[0], flowerScale = p[1], flowerFreq = p[2], flowerPhase = p[3]*time, speed = p[4], roll = p[5]*time, pitch = p[6]*time;
	const float radius = p[7], uMul = p[8], vMul = p[9], fogs[2] = { p[10], p[11] };
	time *= speed;
	for (int iY = 0; iY < s_fxY; ++iY)
		for (int iX = 0; iX < s_fxX; ++iX)
		{
			float u, v; to_uv(iX, iY, 2.f, &u, &v);
			vec4 dir = { u, v, 1.f, 0.f };
			rotX_(pitch, &dir.y, &dir.z);
			rotYZ_(roll, &dir.x, &dir.y);
			fast_norm(&dir);
			float A = dir.x*dir.x + dir.y*dir.y;
			A += flowerScale*orc_lutcosf(atan2f(dir.y, dir.x)*flowerFreq + flowerPhase);
			const float absX = fabsf(dir.x), absY = fabsf(dir.y), box = absX > absY ? absX : absY;
			A = smoothstepf_(A, box, boxy);
			A += KEPSILON;
			A = 1.f/A;
			const float T = radius*A, T2 = T*0.912f;
			const float U = atan2f(dir.y*T, dir.x*T)/KPI, V = dir.z*T + time*speed;
			const float U2 = atan2f(dir.y*T2, dir.x*T2)/KPI, V2 = dir.z*T2 + time*speed;
			const int fpU = ftofp24(U*uMul), fpV = ftofp24(V*vMul), fpU2 = ftofp24(U2*uMul), fpV2 = ftofp24(V2*vMul);
			const float shade = clampf_(0.f, 1.f, 1.f-expf(-0.006f*T*T));
			texc t = prep_uvs(fpU, fpV, 1023, 10);
			const uint32_t c0 = bsamp32(tex, t.i00, t.i10, t.i01, t.i11, t.fu, t.fv);
			t = prep_uvs(fpU2, fpV2, 1023, 10);
			const uint32_t c1 = bsamp32(texGlow, t.i00, t.i10, t.i01, t.i11, t.fu, t.fv);
			uint32_t px = 0, gpx = 0;
			for (int i = 0; i < 4; ++i)
			{
				px |= to_chan_noconv(vlerp((float)((c0 >> (8*i)) & 0xff), fogs[0], shade)) << (8*i);
				gpx |= to_chan_noconv(vlerp((float)((c1 >> (8*i)) & 0xff), fogs[1], shade)) << (8*i);
			}
			pDest[iY*s_fxX + iX] = px;
			pGlowDest[iY*s_fxX + iX] = gpx;
		}
}

/* ------------------------------------------------------------------------------------------------------------------
 * voxel casters: cspanISSE16 and the four ray walkers, one step at a time like the reference
 * ---------------------------------------------------------------------------------------------------------------- */

typedef struct { int c[4]; } col16; /* unpacked 16-bit lanes B,G,R,A */
static col16 unpack16(uint32_t px) { col16 r = { { (int)(px & 0xff), (int)((px >> 8) & 0xff), (int)((px >> 16) & 0xff), (int)(px >> 24) } }; return r; }
static int madd16(int a, int b) { return (int)(int16_t)(a & 0xffff)*(int)(int16_t)(b & 0xffff) + (int)(int16_t)(a >> 16)*(int)(int16_t)(b >> 16); } /* pmaddwd lane */

/* cspanISSE16, cspan.h:47-78; writes are clipped to [lo, hi) (the reference has no clipping: see SURVEY App. B H2) */
static void cspan16(uint32_t *pDest, long pos, int destIncr, long lo, long hi, unsigned length, unsigned drawLength, col16 A, col16 B)
{
	const int divisor = (int)(65536u/length);
	const unsigned preSteps = length - drawLength;
	const int64_t prod = (int64_t)divisor*(int64_t)(int)preSteps; /* _mm_mul_epi32: lanes 0 and 2 */
	const int prodLo = (int)(prod & 0xffffffff), prodHi = (int)(prod >> 32);
	uint32_t from[4]; int step[4];
	for (int i = 0; i < 4; ++i)
	{
		const int delta = B.c[i] - A.c[i];
		step[i] = madd16(delta, divisor);
		from[i] = ((uint32_t)A.c[i] << 16) + (uint32_t)madd16(delta, (i & 1) ? prodHi : prodLo);
	}
	while (drawLength--)
	{
		uint32_t px = 0;
		for (int i = 0; i < 4; ++i)
		{
			const uint32_t v = from[i] >> 16;
			px |= ((v > 32767u) ? 0u : (v > 255u ? 255u : v)) << (8*i);
			from[i] += (uint32_t)step[i];
		}
		if (pos >= lo && pos < hi) pDest[pos] = px;
		pos += destIncr;
	}
}
void orc_cspan16(uint32_t *pDest, int destIncr, unsigned length, unsigned drawLength, uint32_t A, uint32_t B)
{
	cspan16(pDest, 0, destIncr, -(1L<<40), 1L<<40, length, drawLength, unpack16(A), unpack16(B));
}

static int subs16i(int a, int b) { return a > b ? a - b : 0; }
static int adds16i(int a, int b) { return a + b > 65535 ? 65535 : a + b; }

/* Landscape_Draw without the optional TapeWarp32 (landscape.cpp:56-192, 228-240); gamepad state zero */
void orc_landscape(uint32_t *pDest, const uint8_t *heightMap, const uint32_t *colorMap, const uint32_t *fogGradient, float forward, float tiltTrack)
{
	const size_t n = (size_t)s_resX*s_resY;
	for (size_t i = 0; i < n; ++i) pDest[i] = fogGradient[0];
	const float tilt = clampf_(-90.f, 90.f, tiltTrack + 0.f);
	const int mapTilt = 90 + cvtt(tilt);
	const float viewCos = cosf(0.f), viewSin = sinf(0.f);
	const float X1 = -viewSin*forward + 0.f + 0.f, Y1 = viewCos*forward + 0.f + 0.f;
	const int fpX1 = ftofp24(X1), fpY1 = ftofp24(Y1);
	const float rayY = 1024*(s_aspect*(KPI*0.1f));
	for (unsigned iRay = 0; iRay < (unsigned)s_resX; ++iRay)
	{
		const float rayX = 0.25f*(iRay - s_resX*0.5f);
		const float rotRayX = viewCos*rayX - viewSin*rayY, rotRayY = viewSin*rayX + viewCos*rayY; /* vrot2D */
		const float X2 = X1+rotRayX, Y2 = Y1+rotRayY;
		float dX = X2-X1, dY = Y2-Y1;
		if (fabsf(dX+dY) > KEPSILON) { const float length = 1.f/sqrtf(dX*dX + dY*dY); dX *= length; dY *= length; } /* vnorm2D */
		const float fishMul = rayY / sqrtf(rotRayX*rotRayX + rotRayY*rotRayY);
		/* vscape_ray */
		int curX = fpX1, curY = fpY1;
		const int fdX = ftofp24(dX), fdY = ftofp24(dY), fpFishMul = ftofp24(fabsf(fishMul));
		int lastHeight = s_resY, lastDrawnHeight = s_resY;
		col16 lastColor = unpack16(colorMap[(((unsigned)(curX>>8)) & 1023u) | ((((unsigned)(curY>>8)) & 1023u) << 10)]);
		for (unsigned iStep = 0; iStep < 512; ++iStep)
		{
			curX = (int)((unsigned)curX + (unsigned)fdX); curY = (int)((unsigned)curY + (unsigned)fdY);
			const texc t = prep_uvs(curX, curY, 1023, 10);
			const unsigned mapHeight = bsamp8(heightMap, t.i00, t.i10, t.i01, t.i11, (int)t.fu, (int)t.fv);
			col16 color = unpack16(bsamp32(colorMap, t.i00, t.i10, t.i01, t.i11, t.fu, t.fv));
			const col16 fog = unpack16(fogGradient[iStep>>1]);
			for (int i = 0; i < 4; ++i) color.c[i] = subs16i(color.c[i], fog.c[i]);
			int height = 255-(int)mapHeight;
			height <<= 16;
			height = (int)((unsigned)height / ((unsigned)fpFishMul*(iStep+1)));
			height *= 512;
			height >>= 8;
			height += mapTilt;
			if (height < lastDrawnHeight)
			{
				cspan16(pDest + iRay, (long)height*s_resX, s_resX, 0, (long)n, (unsigned)(lastHeight - height), (unsigned)(lastDrawnHeight - height), color, lastColor);
				lastDrawnHeight = height;
			}
			lastHeight = height;
			lastColor = color;
		}
	}
}

/* tscape into the render target (tunnelscape.cpp:44-134, 170-171) */
void orc_tunnelscape_rt(uint32_t *pRT, const uint8_t *heightMap, const uint32_t *colorMap, const uint32_t *fogGradient, float time, float stepU, float stepV, float speed)
{
	const size_t n = (size_t)s_resX*s_resY;
	for (size_t i = 0; i < n; ++i) pRT[i] = fogGradient[0];
	const float mapStepX = 2048.f/(s_resY-1);
	const float speedMul = sqrtf(stepU*stepU + stepV*stepV) * speed;
	const float fromY = 1024.f + speedMul*time;
	const int dX = ftofp24(stepV), dY = ftofp24(s_oneOverAspect*stepU), fpFromY = ftofp24(fromY);
	const float viewLenScale = s_aspect*0.5f;
	for (unsigned iRay = 0; iRay < (unsigned)s_resY; ++iRay)
	{
		const float mapX = iRay*mapStepX, fromX = mapX + stepU * time*KGOLDENRATIO;
		int curX = ftofp24(fromX), curY = fpFromY;
		uint32_t *row = pRT + (size_t)iRay*s_resX;
		long pos = 0;
		int lastHeight = s_resX, lastDrawnHeight = s_resX;
		col16 lastColor = unpack16(colorMap[(((unsigned)(curX>>8)) & 2047u) | ((((unsigned)(curY>>8)) & 2047u) << 11)]);
		for (unsigned iStep = 0; iStep < 512; ++iStep)
		{
			curX = (int)((unsigned)curX + (unsigned)dX); curY = (int)((unsigned)curY + (unsigned)dY);
			const texc t = prep_uvs(curX, curY, 2047, 11);
			const unsigned mapHeight = bsamp8(heightMap, t.i00, t.i10, t.i01, t.i11, (int)t.fu, (int)t.fv);
			col16 color = unpack16(bsamp32(colorMap, t.i00, t.i10, t.i01, t.i11, t.fu, t.fv));
			const col16 fog = unpack16(fogGradient[iStep>>1]);
			for (int i = 0; i < 4; ++i) color.c[i] = subs16i(color.c[i], fog.c[i]);
			int height = 255-(int)mapHeight;
			height -= 96;
			height <<= 8;
			height = cvtt((float)height/viewLenScale);
			height = (int)((unsigned)height / (iStep+1));
			height = (int)((unsigned)height*160u);
			height >>= 8;
			height += 120;
			if (height < lastDrawnHeight)
			{
				const unsigned drawLength = (unsigned)(lastDrawnHeight - height);
				cspan16(row, pos, 1, 0, s_resX, (unsigned)(lastHeight - height), drawLength, color, lastColor);
				lastDrawnHeight = height;
				pos += drawLength;
			}
			lastHeight = height;
			lastColor = color;
		}
	}
}

/* vball into the render target (ball.cpp:80-365).  tables as vball_precalc builds them, computed here with the host libm.
 * ip[]: rayLength, beamAtten, lowLight, hasBeams; fp[]: radius, beamAlphaMin, time (already * ball:Speed), rotateOffsX, rotateOffsY */
void orc_ball_rt(uint32_t *pRT, const uint8_t *heightMix, const uint32_t *colorMap, const uint32_t *auxMap, const int *ip, const float *fp)
{
	const unsigned rayLength = (unsigned)ip[0], beamAtten = (unsigned)ip[1], lowLight = (unsigned)ip[2];
	const int hasBeams = ip[3];
	const float radius = fp[0], beamAlphaMin = fp[1], time = fp[2];
	static unsigned heightProj[1024];
	static int projNorm[1024][3];
	const float angStepSin = KPI/(rayLength-1), angStepCos = angStepSin*0.99f;
	for (unsigned iAngle = 0; iAngle < rayLength; ++iAngle) /* vball_precalc, ball.cpp:283-308 */
	{
		heightProj[iAngle] = f2u(radius*sinf(angStepSin*iAngle));
		const float cosine = cosf(angStepCos*iAngle);
		if (cosine >= 0.f)
		{
			projNorm[iAngle][0] = cvtt(255.f*powf(cosine, KGOLDENRATIO));
			projNorm[iAngle][1] = cvtt(255.f*powf(cosine, KGOLDENANGLE));
			projNorm[iAngle][2] = cvtt(255.f*powf(cosine, KPI));
		}
		else projNorm[iAngle][0] = projNorm[iAngle][1] = projNorm[iAngle][2] = 0;
	}
	const float timeScale = rayLength*(0.25f/1024), fMapDim = 1024.f, fMapHalf = fMapDim*0.5f;
	const int fromX = ftofp24(fMapDim*sinf(time*timeScale) + fMapHalf + fp[3]);
	const int fromY = ftofp24(fMapDim*cosf(time*timeScale) + fMapHalf + fp[4]);
	const float delta = K2PI/(s_resY-1);
	if (!hasBeams) memset(pRT, 0, (size_t)s_resX*s_resY*4);

	for (unsigned iRay = 0; iRay < (unsigned)s_resY; ++iRay)
	{
		const float curAngle = iRay*delta;
		float fdX = cosf(curAngle), fdY = sinf(curAngle); /* calc_fandeltas, voxel-shared.h:26-31 */
		if (fabsf(fdX+fdY) > KEPSILON) { const float length = 1.f/sqrtf(fdX*fdX + fdY*fdY); fdX *= length; fdY *= length; }
		const int dX = ftofp24(fdX), dY = ftofp24(fdY);
		uint32_t *row = pRT + (size_t)iRay*s_resX;
		int curX = fromX, curY = fromY;
		int envU = (1024>>1)<<8, envV = envU;
		unsigned lastHeight = 0, lastDrawnHeight = 0;
		texc t = prep_uvs(curX, curY, 1023, 10);
		col16 lastColor = unpack16(bsamp32(colorMap, t.i00, t.i10, t.i01, t.i11, t.fu, t.fv));
		int beamAccum[4] = { 0, 0, 0, 0 };
		for (unsigned iStep = 0; iStep < rayLength; ++iStep)
		{
			curX = (int)((unsigned)curX - (unsigned)dX); curY = (int)((unsigned)curY - (unsigned)dY);
			t = prep_uvs(curX, curY, 1023, 10);
			const unsigned mapHeight = bsamp8(heightMix, t.i00, t.i10, t.i01, t.i11, (int)t.fu, (int)t.fv);
			col16 color = unpack16(bsamp32(colorMap, t.i00, t.i10, t.i01, t.i11, t.fu, t.fv));
			if (hasBeams) /* vball_ray_beams, ball.cpp:113-140 */
			{
				const col16 beam = unpack16(bsamp32(auxMap, t.i00, t.i10, t.i01, t.i11, t.fu, t.fv));
				const unsigned heightNorm = (mapHeight*(unsigned)projNorm[iStep][0]) >> 8, heightNorm2 = (mapHeight*(unsigned)projNorm[iStep][1]) >> 8;
				const unsigned diffuse = heightNorm + (((unsigned)(int)(heightNorm2-heightNorm)*lowLight) >> 8);
				const int litWhite = (int)(diffuse & 0xffff);
				for (int i = 0; i < 4; ++i)
				{
					const int b = ((beam.c[i]*(int)beamAtten) & 0xffff) >> 8;
					beamAccum[i] = adds16i(beamAccum[i], ((b*litWhite) & 0xffff) >> 8);
					color.c[i] = adds16i(adds16i(color.c[i], beamAccum[i]), litWhite);
				}
			}
			else /* vball_ray_no_beams, ball.cpp:228-262 */
			{
				envU = (int)((unsigned)envU - (unsigned)(dX<<1)); envV = (int)((unsigned)envV - (unsigned)(dY<<1));
				const texc te = prep_uvs(envU + (int)mapHeight, envV + (int)mapHeight, 1023, 10);
				const col16 envCol = unpack16(bsamp32(auxMap, te.i00, te.i10, te.i01, te.i11, te.fu, te.fv));
				const unsigned diffuse = (mapHeight*(unsigned)projNorm[iStep][2]) >> 8;
				const int lit = (int)(diffuse & 0xffff), litFull = (int)(32+diffuse > 255 ? 255 : 32+diffuse);
				for (int i = 0; i < 4; ++i) color.c[i] = adds16i(adds16i(color.c[i], ((envCol.c[i]*litFull) & 0xffff) >> 8), lit);
			}
			const unsigned height = (mapHeight*heightProj[iStep]) >> 8;
			if (height > lastDrawnHeight)
			{
				cspan16(row, lastDrawnHeight, 1, 0, s_resX, height - lastHeight, height - lastDrawnHeight, lastColor, color);
				lastDrawnHeight = height;
			}
			lastHeight = height;
			lastColor = color;
		}
		if (hasBeams) /* beam extrusion, ball.cpp:168-203 */
		{
			uint32_t beamCol = 0;
			for (int i = 0; i < 4; ++i) { const uint32_t v = (uint32_t)beamAccum[i]; beamCol |= ((v > 32767u) ? 0u : (v > 255u ? 255u : v)) << (8*i); }
			const unsigned remainder = (unsigned)(s_resX - 1) - lastDrawnHeight;
			beamCol &= 0xffffff;
			const unsigned beamR = beamCol >> 16, beamG = (beamCol >> 8) & 0xff, beamB = beamCol & 0xff;
			const unsigned mulR = (unsigned)(0.0722f*65536.f), mulG = (unsigned)(0.7152f*65536.f), mulB = (unsigned)(0.2126f*65536.f);
			const float fLuminosity = (float)(((beamR*mulR) >> 16) + ((beamG*mulG) >> 16) + ((beamB*mulB) >> 16));
			const float alphaStep = 1.f / (remainder - 1);
			float curStep = 0.f;
			if (remainder <= (unsigned)s_resX)
				for (unsigned iPixel = 0; iPixel < remainder; ++iPixel)
				{
					const unsigned beamAlpha = f2u(smoothstepf_(beamAlphaMin, fLuminosity, curStep));
					row[lastDrawnHeight++] = beamCol | (beamAlpha << 24);
					curStep += alphaStep;
				}
		}
	}
}

/* vtwister into the render target (torus-twister.cpp:40-135, 169-170) */
void orc_twister_rt(uint32_t *pRT, const uint8_t *heightMap, const uint32_t *colorMap, float time, float speed, float shearSpeed)
{
	static unsigned heightProj[512], heightProjNorm[512];
	for (unsigned iAngle = 0; iAngle < 512; ++iAngle) /* vtwister_precalc */
	{
		const float angle = KPI/(512-1) * iAngle;
		heightProj[iAngle] = f2u(600.f*sinf(angle));
		const float cosine = cosf(angle*0.99f);
		heightProjNorm[iAngle] = (cosine > 0.f) ? f2u(255.f*powf(cosine, 2.f)) : 0;
	}
	memset(pRT, 0, (size_t)s_resX*s_resY*4);
	const float fMapSize = 1024.f, fMapSizeHH = (fMapSize*0.5f) - 0.5f, fMapSizeHHH = (fMapSize*0.25f) - 0.5f;
	const float mapStepY = fMapSize/(s_resY-1);
	for (unsigned iRay = 0; iRay < (unsigned)s_resY; ++iRay)
	{
		const float shearAngle = (float) iRay * (K2PI/(s_resY-1));
		const float mapY = iRay*mapStepY;
		const int fromX = ftofp24(fMapSizeHH + fMapSizeHHH*sinf(time*shearSpeed + shearAngle));
		const int fromY = ftofp24(mapY + time*speed);
		uint32_t *row = pRT + (size_t)iRay*s_resX;
		for (int side = 0; side < 2; ++side)
		{
			int curX = side ? (int)((unsigned)fromX - 512u) : fromX;
			const int dX = side ? -512 : 512, direction = (dX < 0) ? -1 : 1;
			long pos = (s_resX>>1) - side;
			unsigned lastHeight = 0, lastDrawnHeight = 0;
			texc t = prep_uvs(curX, fromY, 1023, 10);
			col16 lastColor = unpack16(bsamp32(colorMap, t.i00, t.i10, t.i01, t.i11, t.fu, t.fv));
			for (unsigned iStep = 0; iStep < 512; ++iStep)
			{
				curX = (int)((unsigned)curX - (unsigned)dX);
				t = prep_uvs(curX, fromY, 1023, 10);
				const unsigned mapHeight = bsamp8(heightMap, t.i00, t.i10, t.i01, t.i11, (int)t.fu, (int)t.fv);
				col16 color = unpack16(bsamp32(colorMap, t.i00, t.i10, t.i01, t.i11, t.fu, t.fv));
				const int litWhite = (int)(((mapHeight*heightProjNorm[iStep]) >> 8) & 0xffff);
				for (int i = 0; i < 4; ++i) color.c[i] = adds16i(color.c[i], litWhite);
				const unsigned height = (mapHeight*heightProj[iStep]) >> 8;
				if (height > lastDrawnHeight)
				{
					const unsigned drawLength = height - lastDrawnHeight;
					cspan16(row, pos, direction, 0, s_resX, height - lastHeight, drawLength, lastColor, color);
					pos += (long)drawLength*direction;
					lastDrawnHeight = height;
				}
				lastHeight = height;
				lastColor = color;
			}
		}
	}
}
