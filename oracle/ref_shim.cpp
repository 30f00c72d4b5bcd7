// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Headless harness that is compiled *together with* the reference's own hot-path translation units
// (taken where they lie under /root/reference/code by oracle/build_ref.py) to produce
// oracle/_ref/libckd_ref_<resY>.so.  It supplies the six symbols the reference expects from its
// SDL/BASS/DevIL services (SURVEY.md section 8c "Link surface") and a small extern "C" surface so the
// Python tests and bench.py can drive the reference's public entry points.
//
// Stubs follow: code/image.h:10-11 (Image_Load32/8), code/gamepad.h:24 (Gamepad_Update),
// code/audio.h:20,27,31 (Audio_Start_Stream, Audio_Rocket_IsPlaying, Audio_Rocket_Sync),
// code/main.h:49 (SetLastError).

#include "main.h"
#include "image.h"
#include "gamepad.h"
#include "rocket.h"
#include "bilinear.h"
#include "cspan.h"
#include "polar.h"
#include "boxblur.h"
#include "deprecated/boxblur.h"
#include "fx-blitter.h"
#include "shadertoy.h"
// sse_mathfun.h defines its functions non-inline (already emitted by shadertoy.cpp): give this TU private copies
#define log_ps shim_log_ps
#define exp_ps shim_exp_ps
#define sin_ps shim_sin_ps
#define cos_ps shim_cos_ps
#define sincos_ps shim_sincos_ps
#include "shadertoy-util.h"
#include "landscape.h"
#include "tunnelscape.h"
#include "ball.h"
#include "torus-twister.h"
#include "demo.h"

#include <unistd.h>
#include <string.h>
#include <map>
#include <string>

// ---------------------------------------------------------------------------------------------
// service stubs
// ---------------------------------------------------------------------------------------------

static std::string s_lastError;
void SetLastError(const std::string &description) { s_lastError = description; }

static double s_timeSec = 0.0;
static const double kRowRateStub = (170.0 / (60.0*(170.0/174.0)))*16.0; // code/audio.cpp:18

void Audio_Start_Stream(unsigned) {}
int Audio_Rocket_IsPlaying(void *) { return 1; }
double Audio_Rocket_Sync(unsigned int &modOrder, unsigned int &modRow, float &modRowAlpha)
{
	modOrder = 0; modRow = 0; modRowAlpha = 0.f;
	return s_timeSec*kRowRateStub; // code/audio.cpp:175-178
}

bool Gamepad_Update(PadState &state)
{
	memset(&state, 0, sizeof(state));
	return false;
}

struct RegImage { const void *pData; size_t numBytes; };
static std::map<std::string, RegImage> s_images;

static void *LoadRegistered(const std::string &path)
{
	auto it = s_images.find(path);
	if (it == s_images.end())
	{
		SetLastError("Can not load image: " + path);
		return nullptr;
	}

	// slack: the reference's samplers may touch a little past the end
	void *pCopy = mallocAligned(it->second.numBytes + 256, kAlignTo);
	memset(pCopy, 0, it->second.numBytes + 256);
	memcpy(pCopy, it->second.pData, it->second.numBytes);
	return pCopy;
}

uint32_t *Image_Load32(const std::string &path) { return static_cast<uint32_t *>(LoadRegistered(path)); }
uint8_t *Image_Load8(const std::string &path) { return static_cast<uint8_t *>(LoadRegistered(path)); }

// ---------------------------------------------------------------------------------------------
// extern "C" surface
// ---------------------------------------------------------------------------------------------

extern "C" {

int ref_res_x() { return int(kResX); }
int ref_res_y() { return int(kResY); }
int ref_fxmap_res_x() { return int(kFxMapResX); }
int ref_fxmap_res_y() { return int(kFxMapResY); }
const char *ref_last_error() { return s_lastError.c_str(); }

// images are decoded by the caller (shared bytes with the CUDA side); the data must stay alive until ref_create()
void ref_register_image(const char *path, const void *pData, size_t numBytes)
{
	s_images[path] = { pData, numBytes };
}

static bool s_withDemo = false;

// baseDir must contain "sync/" with the binary Rocket tracks (code/rocket.cpp:30)
int ref_create(const char *baseDir)
{
	char cwd[4096];
	if (nullptr == getcwd(cwd, sizeof(cwd)))
		return -1;
	if (0 != chdir(baseDir))
		return -2;

	int result = 0;

	// init. order: code/main.cpp:263-279, code/demo.cpp:140-148
	CalculateCosLUT();
	InitializeFastCosine();

	if (!Shared_Create()) result = -3;
	if (0 == result && !Polar_Create()) result = -4;
	if (0 == result && !FxBlitter_Create()) result = -5;
	if (0 == result && !BoxBlur_Create()) result = -6;
	if (s_withDemo)
	{
		// Demo_Create = Rocket::Launch + the five X_Create + the compositor's tracks and art (code/demo.cpp:138-374)
		if (0 == result && !Demo_Create()) result = -14;
	}
	else
	{
		if (0 == result && !Rocket::Launch()) result = -7;
		if (0 == result && !Twister_Create()) result = -8;
		if (0 == result && !Landscape_Create()) result = -9;
		if (0 == result && !Ball_Create()) result = -10;
		if (0 == result && !Tunnelscape_Create()) result = -11;
		if (0 == result && !Shadertoy_Create()) result = -12;
	}

	if (0 != chdir(cwd))
		return -13;

	return result;
}

// the same with the compositor (needs every image of code/demo.cpp:198-374 registered)
int ref_create_demo(const char *baseDir)
{
	s_withDemo = true;
	return ref_create(baseDir);
}

// Demo_Draw, code/demo.cpp:469-1023: runs Rocket::Boost() itself; returns 0 when the demo is over
int ref_demo_draw(uint32_t *pDest, float time, float delta)
{
	return Demo_Draw(pDest, time, delta) ? 1 : 0;
}

void ref_set_time(double seconds)
{
	s_timeSec = seconds;
	Rocket::Boost();
}

double ref_row() { return s_timeSec*kRowRateStub; }

// value of a named track at the current row (tracks are created on first use, code/rocket.cpp:79-82)
double ref_track(const char *baseDir, const char *name)
{
	char cwd[4096];
	if (nullptr == getcwd(cwd, sizeof(cwd)) || 0 != chdir(baseDir))
		return 0.0;
	const sync_track *track = Rocket::AddTrack(name);
	const double value = Rocket::get(track);
	if (0 != chdir(cwd)) return 0.0;
	return value;
}

enum RefEffect
{
	kRefPlasma = 0,
	kRefNautilus = 1,
	kRefSpikeyClose = 2,
	kRefSpikeyDistant = 3,
	kRefTunnel = 4,
	kRefSinuses = 5,
	kRefLaura = 6,
	kRefLandscape = 7,
	kRefTunnelscape = 8,
	kRefBall = 9,
	kRefTwister = 10
};

int ref_draw(int effect, uint32_t *pDest, float time, float delta)
{
	switch (effect)
	{
	case kRefPlasma:        Plasma_Draw(pDest, time, delta); break;
	case kRefNautilus:      Nautilus_Draw(pDest, time, delta); break;
	case kRefSpikeyClose:   Spikey_Draw(pDest, time, delta, true); break;
	case kRefSpikeyDistant: Spikey_Draw(pDest, time, delta, false); break;
	case kRefTunnel:        Tunnel_Draw(pDest, time, delta); break;
	case kRefSinuses:       Sinuses_Draw(pDest, time, delta); break;
	case kRefLaura:         Laura_Draw(pDest, time, delta); break;
	case kRefLandscape:     Landscape_Draw(pDest, time, delta); break;
	case kRefTunnelscape:   Tunnelscape_Draw(pDest, time, delta); break;
	case kRefBall:          Ball_Draw(pDest, time, delta); break;
	case kRefTwister:       Twister_Draw(pDest, time, delta); break;
	default: return -1;
	}
	return 0;
}

uint32_t *ref_fxmap(int index) { return g_pFxMap[index]; }
uint32_t *ref_render_target(int index) { return g_renderTarget[index]; }
const float *ref_cos_lut() { return g_cosLUT; }
const double *ref_fast_cos_tab() { return g_fastCosTab; }

// --- 2D post chain -------------------------------------------------------------------------------

void ref_fx_blit_2x2(uint32_t *pDest, const uint32_t *pSrc) { Fx_Blit_2x2(pDest, pSrc); }
void ref_polar_blit(uint32_t *pDest, const uint32_t *pSrc, int inverse) { Polar_Blit(pDest, pSrc, 0 != inverse); }
void ref_polar_blit_a(uint32_t *pDest, const uint32_t *pSrc, int inverse) { Polar_BlitA(pDest, pSrc, 0 != inverse); }
void ref_polar_blit_2x2(uint32_t *pDest, const uint32_t *pSrc, int inverse) { Polar_Blit_2x2(pDest, pSrc, 0 != inverse); } // needs P2/P2b (build_ref.py)
void ref_fx_test_pattern(uint32_t *pDest) { FxBlitter_DrawTestPattern(pDest); }
uint32_t *ref_ball_background() { return Ball_GetBackground(); }

void ref_old_blur_h(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength) { HorizontalBoxBlur32(pDest, pSrc, xRes, yRes, strength); }
void ref_old_blur_v(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength) { VerticalBoxBlur32(pDest, pSrc, xRes, yRes, strength); }
void ref_old_blur(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength) { BoxBlur32(pDest, pSrc, xRes, yRes, strength); }
float ref_box_blur_scale(float strength) { return BoxBlurScale(strength); }

void ref_new_blur_h(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float gain, unsigned numPasses) { BoxBlur_Horz32(pDest, pSrc, xRes, yRes, strength, gain, numPasses); }
void ref_new_blur_v(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float gain, unsigned numPasses) { BoxBlur_Vert32(pDest, pSrc, xRes, yRes, strength, gain, numPasses); }
void ref_new_blur(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float gain, unsigned numPasses) { BoxBlur_32(pDest, pSrc, xRes, yRes, strength, gain, numPasses); }

void ref_memset32(uint32_t *pDest, int value, size_t numInts) { memset32(pDest, value, numInts); }
void ref_tape_warp(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float speed) { TapeWarp32(pDest, pSrc, xRes, yRes, strength, speed); }

// blend ops (code/util.h:73-122); op ids are shared with include/ckd.h (ckd_blend_op)
int ref_blend(int op, uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels, float fParam, unsigned uParam)
{
	switch (op)
	{
	case 0:  Mix32(pDest, pSrc, numPixels, uint8_t(uParam)); break;
	case 1:  MixOver32(pDest, pSrc, numPixels); break;
	case 2:  Add32(pDest, pSrc, numPixels); break;
	case 3:  Sub32(pDest, pSrc, numPixels); break;
	case 4:  Excl32(pDest, pSrc, numPixels); break;
	case 5:  SoftLight32(pDest, pSrc, numPixels); break;
	case 6:  SoftLight32A(pDest, pSrc, numPixels); break;
	case 7:  SoftLight32AA(pDest, pSrc, numPixels, fParam); break;
	case 8:  Overlay32(pDest, pSrc, numPixels); break;
	case 9:  Overlay32A(pDest, pSrc, numPixels); break;
	case 10: Darken32_50(pDest, pSrc, numPixels); break;
	case 11: MulSrc32(pDest, pSrc, numPixels); break;
	case 12: MulSrc32A(pDest, pSrc, numPixels); break;
	case 13: MixSrc32(pDest, pSrc, numPixels); break;
	case 14: Fade32(pDest, numPixels, uParam & 0xffffff, uint8_t(uParam >> 24)); break;
	default: return -1;
	}
	return 0;
}

// rectangular blits (code/util.h:103-117)
int ref_blit(int op, uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned srcResX, unsigned yRes, float alpha)
{
	switch (op)
	{
	case 0: BlitSrc32(pDest, pSrc, destResX, srcResX, yRes); break;
	case 1: BlitSrc32A(pDest, pSrc, destResX, srcResX, yRes, alpha); break;
	case 2: BlitAdd32(pDest, pSrc, destResX, srcResX, yRes); break;
	case 3: BlitAdd32A(pDest, pSrc, destResX, srcResX, yRes, alpha); break;
	default: return -1;
	}
	return 0;
}

void ref_mix_src_s(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned destResY, unsigned srcStride) { MixSrc32S(pDest, pSrc, destResX, destResY, srcStride); }

// --- scalar probes for unit-level parity of the device math layer --------------------------------

float ref_lutcosf(float x) { return lutcosf(x); }
float ref_lutsinf(float x) { return lutsinf(x); }
float ref_fastcosf(double x) { return fastcosf(x); }
float ref_q3_rsqrtf2(float x) { return Q3_rsqrtf<2>(x); }

float ref_rsqrt_ss(float x)
{
	float result;
	_mm_store_ss(&result, _mm_rsqrt_ss(_mm_set_ss(x)));
	return result;
}

// fills table[parity*2^23 + mantissa] for exponent parities {126 (x in [0.5,1)), 127 (x in [1,2))}, subsampled by 'stride'
void ref_rsqrt_scan(uint32_t *pTable, unsigned stride)
{
	size_t iOut = 0;
	for (unsigned parity = 0; parity < 2; ++parity)
		for (unsigned mant = 0; mant < (1u<<23); mant += stride)
		{
			const uint32_t bits = ((126u+parity)<<23) | mant;
			const float result = ref_rsqrt_ss(std::bit_cast<float>(bits));
			pTable[iOut++] = std::bit_cast<uint32_t>(result);
		}
}

void ref_log_ps(const float *pIn, float *pOut) { _mm_storeu_ps(pOut, log_ps(_mm_loadu_ps(pIn))); }
void ref_exp_ps(const float *pIn, float *pOut) { _mm_storeu_ps(pOut, exp_ps(_mm_loadu_ps(pIn))); }

// GammaAdj + ToPixel4 on n colours of 4 lanes (code/shadertoy-util.h:136-146,186-190); n must be a multiple of 4
void ref_gamma_pixels(const float *pColors, float gamma, uint32_t *pOut, unsigned n)
{
	for (unsigned i = 0; i < n; i += 4)
	{
		__m128 colors[4];
		for (int j = 0; j < 4; ++j)
			colors[j] = Shadertoy::GammaAdj(_mm_loadu_ps(pColors + (i+j)*4), gamma);
		_mm_storeu_si128(reinterpret_cast<__m128i *>(pOut + i), Shadertoy::ToPixel4(colors));
	}
}

void ref_to_pixels_noconv(const float *pColors, uint32_t *pOut, unsigned n)
{
	for (unsigned i = 0; i < n; i += 4)
	{
		__m128 colors[4];
		for (int j = 0; j < 4; ++j)
			colors[j] = _mm_loadu_ps(pColors + (i+j)*4);
		_mm_storeu_si128(reinterpret_cast<__m128i *>(pOut + i), Shadertoy::ToPixel4_NoConv(colors));
	}
}

// host libm probes (the reference calls the host's glibc for these)
float ref_powf(float a, float b) { return powf(a, b); }
float ref_expf(float a) { return expf(a); }
float ref_atan2f(float a, float b) { return atan2f(a, b); }

// cspanISSE16 on packed colours (code/cspan.h:47-78)
void ref_cspan16(uint32_t *pDest, int destIncr, unsigned length, unsigned drawLength, uint32_t A, uint32_t B)
{
	cspanISSE16(pDest, destIncr, length, drawLength, c2vISSE16(A), c2vISSE16(B));
}

// bilinear samplers (code/bilinear.h)
unsigned ref_bsamp8(const uint8_t *pTexture, int U, int V, unsigned mapAnd, unsigned mapShift)
{
	unsigned U0, V0, U1, V1, fracU, fracV;
	bsamp_prepUVs(U, V, mapAnd, mapShift, U0, V0, U1, V1, fracU, fracV);
	return bsamp8(pTexture, U0, V0, U1, V1, fracU, fracV);
}

uint32_t ref_bsamp32(const uint32_t *pTexture, int U, int V, unsigned mapAnd, unsigned mapShift)
{
	unsigned U0, V0, U1, V1, fracU, fracV;
	bsamp_prepUVs(U, V, mapAnd, mapShift, U0, V0, U1, V1, fracU, fracV);
	return v2cISSE16(bsamp32_16(pTexture, U0, V0, U1, V1, fracU, fracV));
}

} // extern "C"
