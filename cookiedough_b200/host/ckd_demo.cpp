// ckd_demo.cpp -- the compositor: Demo_Create / Demo_Draw / Demo_Destroy of the reference (demo.h:8-10, demo.cpp:138-1023).
//
// Demo_Draw picks the part from the "demo:Effect" track, renders its effect and lays the part's art over it with the blend /
// blit / blur / warp operations of util.cpp, deprecated/boxblur.cpp and polar.cpp.  Here the whole chain runs on the device:
// the effect renders into the device frame, every layer is a device-resident image, the scratch buffers are the context's
// render targets (the reference's g_renderTarget[0..3], demo.cpp:393,427,558,785,873), and the finished frame is copied to
// the caller's pDest once.  Every operation of the chain is integer arithmetic that is bit-exact with the reference, so a
// composed frame differs from the reference's only where the effect underneath does.

#include "ckd_host_internal.h"
#include "../csrc/ckd_hostmath.h"

#include <math.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

using ckdhost::Check;

namespace {

// ---- sync tracks, demo.cpp:28-57 -----------------------------------------------------------------------------------------

SyncTrack trackEffect;
SyncTrack trackFadeToBlack, trackFadeToWhite;
SyncTrack trackCreditLogo, trackCreditLogoAlpha, trackCreditLogoBlurH, trackCreditLogoBlurV;
SyncTrack trackDiscoGuys, trackDiscoGuysAppearance[8];
SyncTrack trackCheapJoke;
SyncTrack trackShow1995, trackShow2006;
SyncTrack trackDirt;
SyncTrack trackScapeOverlay, trackScapeRevision, trackScapeFade;
SyncTrack trackDistortTPB, trackDistortStrengthTPB, trackBlurTPB, trackRibbonsTPB;
SyncTrack trackGreetSwitch;
SyncTrack trackCousteau, trackCousteauHorzBlur;
SyncTrack trackShooting, trackShootingX, trackShootingY, trackShootingAlpha, trackShootingTrail;
SyncTrack trackWaterLove, trackLoveBlurHorz;
SyncTrack trackCloseUpMoonraker, trackCloseUpMoonrakerText, trackCloseUpMoonrakerTextBlur;
SyncTrack trackSpikeDemoLogoIndex;
SyncTrack trackCreditLogoBlend;
SyncTrack trackFullWarpTPB;

// ---- device-resident art -------------------------------------------------------------------------------------------------

struct Layer { uint32_t *d = nullptr; int width = 0, height = 0; };
std::vector<void *> s_allocations;

constexpr int kCredX = 1280, kCredY = 568;      // demo.cpp:63-64
constexpr unsigned kLenzSize = 64;              // demo.cpp:133

Layer s_credits[4], s_comatron[5], s_superplek[5], s_jadeNytrik[5], s_ernstHot[5];
Layer s_vignette06;
Layer s_noooN[4], s_mfx[4], s_tunnelFullDirt, s_tunnelVignette, s_tunnelVignette2;
Layer s_spikeyFullDirt, s_spikeyBypass, s_spikeyArrested[4], s_spikeyVignette, s_spikeyVignette2;
Layer s_godLayer, s_revLogo;
Layer s_ballVignette;
Layer s_greetingsDirt, s_greetings[4], s_greetingsVignette;
Layer s_nautilusVignette, s_nautilusDirt, s_nautilusCousteau1, s_nautilusCousteauRim1, s_nautilusCousteauRim2, s_nautilusCousteau2, s_nautilusText;
Layer s_discoGuys[8], s_areWeDone;
Layer s_closeSpikeDirtRaker, s_closeSpikeVignette, s_closeSpikeVignetteForRaker, s_closeSpike1961;
Layer s_waterDirt, s_waterPrismOverlay;
Layer s_lenz, s_ribbons, s_gpuJoke;
Layer s_nytrikTPB, s_xboxLogoTPB;              // shared-resources.cpp:27-34

bool s_created = false;

// Image_Load32 (image.cpp:31-73) for a layer that lives on the device.  minWidth/minHeight: what the compositor reads from it.
bool LoadLayer(Layer &layer, const char *path, int minWidth, int minHeight)
{
	ckd_ctx *ctx = CkdHost_Context();
	ckdhost::ImageView view;
	if (nullptr == ctx || !ckdhost::FindImage(path, view) || 4 != view.bpp)
	{
		SetLastError(std::string("Can not load image: ") + path); // image.cpp:40
		return false;
	}
	if (view.width < minWidth || view.height < minHeight)
	{
		SetLastError(std::string(path) + " is smaller than the area the compositor reads from it");
		return false;
	}

	const size_t bytes = size_t(view.width)*view.height*4;
	void *d = nullptr;
	if (!Check(ckd_malloc(ctx, &d, bytes + 256), path)) // same slack as the reference harness gives Image_Load32
		return false;
	s_allocations.push_back(d);
	if (!Check(ckd_upload(ctx, d, view.pixels, bytes), path) || !Check(ckd_sync(ctx), path))
		return false;
	layer.d = static_cast<uint32_t *>(d);
	layer.width = view.width;
	layer.height = view.height;
	ckdhost::ReleaseImage(path);
	return true;
}

// ---- the frame being composed --------------------------------------------------------------------------------------------

ckd_ctx *s_c = nullptr;
unsigned s_resX = 0, s_resY = 0, s_outputSize = 0;

// Consecutive blends onto the same buffer are recorded and issued as one fused pass (ckd_blend_chain): the pixel stays in
// registers, every layer is read once.  Anything else that touches device memory flushes the recording first, so the
// order of operations the reference performs is kept.
uint32_t *s_chainDest = nullptr;
unsigned s_chainPixels = 0;
std::vector<ckd_blend_step> s_chain;

bool FlushChain()
{
	if (s_chain.empty())
		return true;
	const bool ok = Check(ckd_blend_chain(s_c, s_chainDest, s_chain.data(), unsigned(s_chain.size()), s_chainPixels), "Demo_Draw: blend");
	s_chain.clear();
	return ok;
}

bool Blend(ckd_blend_op op, uint32_t *d_dest, const uint32_t *d_src, unsigned numPixels, float f = 0.f, unsigned u = 0)
{
	bool ok = true;
	if (!s_chain.empty() && (d_dest != s_chainDest || numPixels != s_chainPixels))
		ok = FlushChain();
	// a layer that was itself produced by the pending chain cannot happen: the chain only ever writes its destination
	s_chainDest = d_dest;
	s_chainPixels = numPixels;
	s_chain.push_back({ op, d_src, f, u });
	return ok;
}

// every other device operation: flush the recorded blends, then run it
#define CKD_DIRECT(call, what) (FlushChain(), Check((call), (what)))

bool Blit(ckd_blit_op op, uint32_t *d_dest, const uint32_t *d_src, unsigned destResX, unsigned srcResX, unsigned yRes, float alpha = 1.f)
{
	return CKD_DIRECT(ckd_blit(s_c, op, d_dest, d_src, destResX, srcResX, yRes, alpha), "Demo_Draw: blit");
}
bool Full(ckd_blend_op op, uint32_t *d_dest, const Layer &layer, float f = 0.f) { return Blend(op, d_dest, layer.d, s_outputSize, f); }

// Fade32(pDest, n, RGB, alpha), util.cpp:798-820
bool Fade(uint32_t *d_dest, uint32_t rgb, uint8_t alpha) { return Blend(CKD_FADE32, d_dest, nullptr, s_outputSize, 0.f, (unsigned(alpha) << 24) | (rgb & 0xffffffu)); }

// float -> uint8_t as gcc compiles it on x86-64: cvttss2si, low byte
uint8_t ToU8(float f) { return uint8_t(ckdh::x86_cvtt(f)); }

// demo.cpp:383-390
void FadeFlash(uint32_t *d_dest, float fadeToBlack, float fadeToWhite)
{
	if (fadeToWhite > 0.f)
		Fade(d_dest, 0xffffff, ToU8(fadeToWhite*255.f));
	if (fadeToBlack > 0.f)
		Fade(d_dest, 0, ToU8(fadeToBlack*255.f));
}

// BloodBlend / CreditBlend, demo.cpp:393-467: cross-fade logo N into logo N+1 in g_renderTarget[3]
const uint32_t *LogoBlend(float blend, const Layer *logos, int numLogos, unsigned width, unsigned height)
{
	uint32_t *d_target = ckd_render_target(s_c, 3);
	const float factor = fmodf(blend, 1.f);
	const uint8_t iFactor = ToU8(255.f*factor);
	const int last = numLogos - 1;

	if (blend >= float(last))
		return logos[last].d;
	for (int i = 0; i < last; ++i)
	{
		if (blend >= float(i) && blend < float(i+1))
		{
			CKD_DIRECT(ckd_copy(s_c, d_target, logos[i].d, size_t(width)*height*4), "Demo_Draw: memcpy");
			Blend(CKD_MIX32, d_target, logos[i+1].d, width*height, 0.f, iFactor);
			break;
		}
	}
	return d_target; // (a negative blend returns the target untouched, like the reference)
}

// synth-math-easings.h:138-143, 177-188 and Std3DMath-stripped/Math.h:67-71
float easeInBackf(float x)
{
	constexpr float c1 = 1.70158f;
	constexpr float c3 = c1 + 1.f;
	return c3 * x*x*x - c1*x*x;
}
float easeOutElasticf(float x)
{
	constexpr float c4 = (2.f*ckdh::kPI)/3.f;
	return (0.f == x) ? 0.f : (1.f == x) ? 1.f : powf(2.f, -10.f*x) * sinf((x*10.f - 0.75f) * c4) + 1.f;
}
float smootherstepf(float a, float b, float t)
{
	t = t*t*t*(t*(t*6.f - 15.f) + 10.f);
	return ckdh::lerpf(a, b, t);
}

// FxBlitter_DrawTestPattern, fx-blitter.cpp:77-98
void DrawTestPattern(uint32_t *d_dest)
{
	const unsigned fxX = unsigned(ckd_fxmap_res_x(s_c)), fxY = unsigned(ckd_fxmap_res_y(s_c));
	std::vector<uint32_t> pattern(size_t(fxX)*fxY);
	for (unsigned iY = 0; iY < fxY; ++iY)
		for (unsigned iX = 0; iX < fxX; ++iX)
			pattern[size_t(iY)*fxX + iX] = (iY < fxY/2) ? ((iY & 1) ? 0xffffffffu : 0u) : ((iX & 1) ? 0xffffffffu : 0u);
	if (CKD_DIRECT(ckd_upload(s_c, ckd_fxmap(s_c, 0), pattern.data(), pattern.size()*4), "FxBlitter_DrawTestPattern") && Check(ckd_sync(s_c), "FxBlitter_DrawTestPattern"))
		CKD_DIRECT(ckd_fx_blit_2x2(s_c, d_dest, ckd_fxmap(s_c, 0)), "FxBlitter_DrawTestPattern");
}

} // namespace

// -----------------------------------------------------------------------------------------------------------------------
// Demo_Create, demo.cpp:138-367
// -----------------------------------------------------------------------------------------------------------------------

bool Demo_Create()
{
	ckd_ctx *ctx = CkdHost_Context();
	if (nullptr == ctx)
	{
		SetLastError("CkdHost_Create() has not been called");
		return false;
	}
	const int resX = ckd_res_x(ctx), resY = ckd_res_y(ctx);

	if (false == Rocket::Launch())
		return false;

	bool fxInit = true;
	fxInit &= Twister_Create();
	fxInit &= Landscape_Create();
	fxInit &= Ball_Create();
	fxInit &= Tunnelscape_Create();
	fxInit &= Shadertoy_Create();

	trackEffect = Rocket::AddTrack("demo:Effect");
	trackFadeToBlack = Rocket::AddTrack("demo:FadeToBlack");
	trackFadeToWhite = Rocket::AddTrack("demo:FadeToWhite");
	trackCreditLogo = Rocket::AddTrack("demo:CreditLogo");
	trackCreditLogoAlpha = Rocket::AddTrack("demo:CreditLogoAlpha");
	trackCreditLogoBlurH = Rocket::AddTrack("demo:CreditLogoBlurH");
	trackCreditLogoBlurV = Rocket::AddTrack("demo:CreditLogoBlurV");
	trackDiscoGuys = Rocket::AddTrack("demo:DiscoGuys");
	for (int i = 0; i < 8; ++i)
		trackDiscoGuysAppearance[i] = Rocket::AddTrack(("demo:DiscoGuy" + std::to_string(i+1)).c_str());
	trackShow1995 = Rocket::AddTrack("demo:Show1995");
	trackDirt = Rocket::AddTrack("demo:Dirt");
	trackShow2006 = Rocket::AddTrack("demo:Show2006");
	trackScapeOverlay = Rocket::AddTrack("demo:ScapeOverlay");
	trackScapeRevision = Rocket::AddTrack("demo:ScapeRev");
	trackScapeFade = Rocket::AddTrack("demo:ScapeFade");
	trackDistortTPB = Rocket::AddTrack("demo:DistortTPB");
	trackDistortStrengthTPB = Rocket::AddTrack("demo:DistortStrengthTPB");
	trackBlurTPB = Rocket::AddTrack("demo:BlurTPB");
	trackRibbonsTPB = Rocket::AddTrack("demo:RibbonsX");
	trackGreetSwitch = Rocket::AddTrack("demo:GreetSwitch");
	trackCousteau = Rocket::AddTrack("demo:Cousteau");
	trackCousteauHorzBlur = Rocket::AddTrack("demo:CousteauHorzBlur");
	trackWaterLove = Rocket::AddTrack("demo:WaterLove");
	trackLoveBlurHorz = Rocket::AddTrack("demo:LoveBlurHorZ");
	trackShooting = Rocket::AddTrack("shootingStar:Enabled");
	trackShootingX = Rocket::AddTrack("shootingStar:X");
	trackShootingY = Rocket::AddTrack("shootingStar:Y");
	trackShootingAlpha = Rocket::AddTrack("shootingStar:A");
	trackShootingTrail = Rocket::AddTrack("shootingStar:Trail");
	trackCheapJoke = Rocket::AddTrack("demo:CheapGPU");
	trackSpikeDemoLogoIndex = Rocket::AddTrack("demo:MainLogoIndex");
	trackCreditLogoBlend = Rocket::AddTrack("demo:CreditAnimBlend");
	trackFullWarpTPB = Rocket::AddTrack("demo:FullWarpTPB");
	trackCloseUpMoonraker = Rocket::AddTrack("closeSpike:Moonraker");
	trackCloseUpMoonrakerText = Rocket::AddTrack("closeSpike:MoonrakerText");
	trackCloseUpMoonrakerTextBlur = Rocket::AddTrack("closeSpike:MoonrakerBlur");

	struct Entry { Layer *layer; std::string path; int width, height; };
	std::vector<Entry> entries;
	auto credits = [&](Layer *l, const std::string &path) { entries.push_back({ l, path, kCredX, kCredY }); };
	auto full = [&](Layer *l, const std::string &path) { entries.push_back({ l, path, resX, resY }); };
	auto sprite = [&](Layer *l, const std::string &path, int w, int h) { entries.push_back({ l, path, w, h }); };

	// Shared_Create, shared-resources.cpp:27-34
	full(&s_nytrikTPB, "assets/demo/TPB-logo.png");
	sprite(&s_xboxLogoTPB, "assets/demo/tpb_xbox_tp-263x243.png", 263, 243);

	credits(&s_credits[0], "assets/credits/Credits_Tag_Superplek_outlined.png");
	credits(&s_credits[1], "assets/credits/Credits_Tag_Comatron_Featuring_Celin_outlined.png");
	credits(&s_credits[2], "assets/credits/Credits_Tag_Jade_outlined.png");
	credits(&s_credits[3], "assets/credits/Credits_Tag_ErnstHot_outlined_new.png");
	for (int i = 0; i < 5; ++i)
	{
		credits(&s_comatron[i], "assets/credits/comatron_anim/comatron_" + std::to_string(i+1) + ".png");
		credits(&s_superplek[i], "assets/credits/animplek/animplek" + std::to_string(i) + ".png");
		credits(&s_jadeNytrik[i], "assets/credits/jade&nytrik/jade&nytrik" + std::to_string(i) + ".png");
		credits(&s_ernstHot[i], "assets/credits/animhot0/animhot" + std::to_string(i) + ".png");
	}

	full(&s_vignette06, "assets/demo/tpb-06-dirty-vignette-1280x720.png");

	for (int i = 0; i < 4; ++i)
	{
		full(&s_spikeyArrested[i], "assets/spikeball/Layer 2023_" + std::to_string(i+1) + ".png");
		full(&s_noooN[i], "assets/tunnels/layer 1995_" + std::to_string(i+1) + ".png");
		full(&s_mfx[i], "assets/tunnels/layer 2006_" + std::to_string(i+1) + ".png");
		full(&s_greetings[i], "assets/greetings/Greetings_Part" + std::to_string(i+1) + "_BG_Overlay.png");
	}
	full(&s_spikeyVignette, "assets/spikeball/Vignette_CoolFilmLook.png");
	full(&s_spikeyVignette2, "assets/spikeball/Vignette_Layer02_inverted.png");
	full(&s_spikeyBypass, "assets/spikeball/SpikeyBall_byPass_BG_Overlay.png");
	full(&s_spikeyFullDirt, "assets/spikeball/nytrik-TheYearWas_Overlay_LensDirt.jpg");
	full(&s_tunnelFullDirt, "assets/tunnels/nytrik-TheYearWas_Overlay_LensDirt.png");
	full(&s_tunnelVignette, "assets/tunnels/Vignette_CoolFilmLook.png");
	full(&s_tunnelVignette2, "assets/tunnels/Vignette_Layer02_inverted.png");
	full(&s_godLayer, "assets/demo/nytrik-god-layer-720p.png");
	full(&s_revLogo, "assets/scape/revision-logo_white.png");
	full(&s_ballVignette, "assets/ball/Vignette_Sparta300.png");
	full(&s_greetingsDirt, "assets/greetings/Bokeh_Lens_Dirt_51.png");
	full(&s_greetingsVignette, "assets/greetings/Vignette_CoolFilmLook.png");
	full(&s_nautilusVignette, "assets/nautilus/Vignette.png");
	full(&s_nautilusDirt, "assets/nautilus/GlassDirt_Distorted2.png");
	full(&s_nautilusCousteau2, "assets/nautilus/JacquesCousteau_Silhouette2.png");
	full(&s_nautilusCousteau1, "assets/nautilus/JacquesCousteau1_Silhouette.png");
	full(&s_nautilusCousteauRim1, "assets/nautilus/JacquesCousteau1_Silhouette_RimMask.png");
	full(&s_nautilusCousteauRim2, "assets/nautilus/JacquesCousteau_Silhouette2_RimMask.png");
	full(&s_nautilusText, "assets/nautilus/JacquesCousteau_Text.png");

	static const char *guys[8] = { "1", "1b", "2", "2b", "3", "3b", "4", "4b" };
	for (int i = 0; i < 8; ++i)
		sprite(&s_discoGuys[i], std::string("assets/demo/tpb-06-disco-guy/") + guys[i] + ".png", 128, 128);
	sprite(&s_areWeDone, "assets/demo/are-we-done-1100x57.png", 1100, 57);

	full(&s_closeSpikeDirtRaker, "assets/closeup/raker-LensDirt5_invert.png");
	full(&s_closeSpikeVignetteForRaker, "assets/closeup/VignetteForRaker.png");
	full(&s_closeSpikeVignette, "assets/closeup/Vignette_CoolFilmLook.png");
	sprite(&s_closeSpike1961, "assets/closeup/raker_textSmall.png", 624, 115);

	full(&s_waterDirt, "assets/underwater/LensDirt3_invert.png");
	full(&s_waterPrismOverlay, "assets/underwater/love prism_alpha 1280_720.png");

	sprite(&s_lenz, "assets/shooting/Lenz.png", kLenzSize, kLenzSize);
	sprite(&s_ribbons, "assets/demo/ribbons.png", 1, 1);
	sprite(&s_gpuJoke, "assets/demo/GPU-joke.png", 960, 160);

	for (const Entry &entry : entries)
		if (!LoadLayer(*entry.layer, entry.path.c_str(), entry.width, entry.height))
			return false;

	// part 12 reads resY - 1 rows of resX pixels, 2160 pixels apart, starting up to resX pixels into the strip (demo.cpp:882-883)
	if (size_t(s_ribbons.width)*s_ribbons.height < size_t(2*resX) + size_t(2160)*(resY - 2))
	{
		SetLastError("assets/demo/ribbons.png is smaller than the area the compositor reads from it");
		return false;
	}

	s_created = fxInit;
	return fxInit;
}

// demo.cpp:370-381
void Demo_Destroy()
{
	Rocket::Land();

	Twister_Destroy();
	Landscape_Destroy();
	Ball_Destroy();
	Tunnelscape_Destroy();
	Shadertoy_Destroy();

	ckd_ctx *ctx = CkdHost_Context();
	if (nullptr != ctx)
		for (void *d : s_allocations)
			ckd_free(ctx, d);
	s_allocations.clear();
	s_created = false;
}

// -----------------------------------------------------------------------------------------------------------------------
// Demo_Draw, demo.cpp:469-1023
// -----------------------------------------------------------------------------------------------------------------------

bool Demo_Draw(uint32_t *pDest, float timer, float delta)
{
	if (!s_created)
	{
		SetLastError("Demo_Create() has not been called");
		return false;
	}

	if (false == Rocket::Boost())
		return false; // demo is over!

	s_c = CkdHost_Context();
	s_resX = unsigned(ckd_res_x(s_c));
	s_resY = unsigned(ckd_res_y(s_c));
	s_outputSize = s_resX*s_resY;
	const unsigned kResX = s_resX, kResY = s_resY, kOutputSize = s_outputSize;

	s_chain.clear();
	uint32_t *d = ckdhost::BeginCompose();       // the device twin of pDest; X_Draw(pDest, ...) renders into it while composing
	uint32_t *rt0 = ckd_render_target(s_c, 0), *rt1 = ckd_render_target(s_c, 1), *rt2 = ckd_render_target(s_c, 2), *rt3 = ckd_render_target(s_c, 3);

	const float fadeToBlack = Rocket::getf(trackFadeToBlack);
	const float fadeToWhite = Rocket::getf(trackFadeToWhite);

	const int effect = Rocket::geti(trackEffect);
	switch (effect)
	{
	case 1: // voxel torus, demo.cpp:511-521
		Twister_Draw(pDest, timer, delta);
		FadeFlash(d, fadeToBlack, fadeToWhite);
		Full(CKD_SOFTLIGHT32A, d, s_closeSpikeVignette);
		Full(CKD_MULSRC32A, d, s_vignette06);
		break;

	case 2: // landscape, demo.cpp:523-592
		{
			Landscape_Draw(pDest, timer, delta);

			const float scapeFade = ckdh::saturatef(Rocket::getf(trackScapeFade));
			FadeFlash(d, scapeFade, 0.f);

			if (1 == Rocket::geti(trackShooting))
			{
				int xPos = Rocket::geti(trackShootingX);
				int yPos = Rocket::geti(trackShootingY);
				float alpha = Rocket::getf(trackShootingAlpha);

				// the reference trusts the tracks to keep the sprite inside the frame (demo.cpp:541-542); so does this, but loudly
				auto lenz = [&](int x, int y, float a)
				{
					if (x < 0 || y < 0 || x + int(kLenzSize) > int(kResX) || y + int(kLenzSize) > int(kResY))
					{
						SetLastError("Demo_Draw: shooting star leaves the frame");
						return;
					}
					Blit(CKD_BLITADD32A, d + unsigned(y)*kResX + unsigned(x), s_lenz.d, kResX, kLenzSize, kLenzSize, a);
				};
				lenz(xPos, yPos, alpha);

				const int trail = Rocket::geti(trackShootingTrail);
				if (trail > 0)
				{
					const int xStep = kLenzSize/16;
					const int yStep = 1;
					const float alphaStep = alpha/trail;
					for (int iTrail = 0; iTrail < trail; ++iTrail)
					{
						xPos += xStep;
						yPos -= yStep;
						alpha -= alphaStep;
						lenz(xPos, yPos, alpha);
					}
				}
			}

			const float overlayAlpha = ckdh::saturatef(Rocket::getf(trackScapeOverlay));
			if (0.f != overlayAlpha)
				Blit(CKD_BLITADD32A, d, s_godLayer.d, kResX, kResX, kResY, overlayAlpha);

			const float alphaRev = ckdh::saturatef(Rocket::getf(trackScapeRevision));
			if (0.f != alphaRev)
			{
				if (alphaRev < 0.314f)
				{
					const float easeA = easeOutElasticf(alphaRev)*ckdh::kGoldenAngle;
					const float easeB = easeInBackf(alphaRev)*ckdh::kGoldenRatio;
					CKD_DIRECT(ckd_tape_warp(s_c, rt0, s_revLogo.d, kResX, kResY, easeA, easeB), "Demo_Draw: TapeWarp32");
				}
				else
					CKD_DIRECT(ckd_old_blur(s_c, rt0, s_revLogo.d, kResX, kResY, ckdh::BoxBlurScale((alphaRev-0.314f)*ckdh::k2PI)), "Demo_Draw: BoxBlur32");
				Blit(CKD_BLITSRC32A, d, rt0, kResX, kResX, kResY, alphaRev);
			}

			FadeFlash(d, fadeToBlack, fadeToWhite);
			Full(CKD_SOFTLIGHT32A, d, s_closeSpikeVignette);
		}
		break;

	case 3: // voxel ball, demo.cpp:594-610
		Ball_Draw(pDest, timer, delta);
		if (false == Ball_HasBeams())
			Full(CKD_MULSRC32, d, s_greetingsVignette);
		else
			Full(CKD_SOFTLIGHT32, d, s_ballVignette);
		FadeFlash(d, fadeToBlack, fadeToWhite);
		if (true == Ball_HasBeams())
			Full(CKD_MULSRC32A, d, s_vignette06);
		break;

	case 4: // tunnels, demo.cpp:612-639
		{
			Tunnelscape_Draw(pDest, timer, delta);
			Full(CKD_SUB32, d, s_tunnelVignette2);
			Full(CKD_MIXSRC32, d, s_tunnelFullDirt);

			const float show1995 = ckdh::clampf(0.f, 3.f, Rocket::getf(trackShow1995));
			if (show1995 > 0.f)
				Blend(CKD_MIXOVER32, d, LogoBlend(show1995, s_noooN, 4, kResX, kResY), kOutputSize);

			Full(CKD_OVERLAY32, d, s_tunnelVignette);
		}
		break;

	case 5: // plasma and credits, demo.cpp:641-718
		{
			Plasma_Draw(pDest, timer, delta);

			const int iLogo = ckdh::clampi(0, 4, Rocket::geti(trackCreditLogo));
			if (0 != iLogo)
			{
				const float logoBlend = ckdh::clampf(0.f, 4.f, Rocket::getf(trackCreditLogoBlend));
				const Layer *logos = (1 == iLogo) ? s_superplek : (2 == iLogo) ? s_comatron : (3 == iLogo) ? s_jadeNytrik : s_ernstHot;

				const uint32_t *pCur = LogoBlend(logoBlend, logos, 5, kCredX, kCredY);

				const float blurH = Rocket::getf(trackCreditLogoBlurH);
				if (0.f != blurH)
				{
					CKD_DIRECT(ckd_old_blur_h(s_c, rt0, pCur, kCredX, kCredY, ckdh::BoxBlurScale(blurH)), "Demo_Draw: HorizontalBoxBlur32");
					pCur = rt0;
				}

				const float blurV = Rocket::getf(trackCreditLogoBlurV);
				if (0 != blurV)
				{
					CKD_DIRECT(ckd_old_blur_v(s_c, rt0, pCur, kCredX, kCredY, ckdh::BoxBlurScale(blurV)), "Demo_Draw: VerticalBoxBlur32");
					pCur = rt0;
				}

				Blit(CKD_BLITSRC32A, d + ((kResY-kCredY)>>1)*kResX, pCur, kResX, kCredX, kCredY, ckdh::clampf(0.f, 1.f, Rocket::getf(trackCreditLogoAlpha)));
			}
		}
		break;

	case 6: // nautilus, demo.cpp:720-760
		{
			Nautilus_Draw(pDest, timer, delta);
			Full(CKD_SOFTLIGHT32, d, s_nautilusVignette);
			Full(CKD_SOFTLIGHT32, d, s_nautilusDirt);
			FadeFlash(d, fadeToBlack, 0.f);

			const bool first = 0 == Rocket::geti(trackCousteau);
			const uint32_t *pCousteau = first ? s_nautilusCousteau1.d : s_nautilusCousteau2.d;
			const Layer &rim = first ? s_nautilusCousteauRim1 : s_nautilusCousteauRim2;

			Full(CKD_OVERLAY32A, d, rim);

			float hBlur = Rocket::getf(trackCousteauHorzBlur);
			if (0.f != hBlur)
			{
				hBlur = ckdh::BoxBlurScale(hBlur);
				CKD_DIRECT(ckd_old_blur_h(s_c, rt0, pCousteau, kResX, kResY, hBlur), "Demo_Draw: HorizontalBoxBlur32");
				pCousteau = rt0;
			}

			Blend(CKD_MIXSRC32, d, pCousteau, kOutputSize);
			FadeFlash(d, 0.f, fadeToWhite);
			Full(CKD_MIXSRC32, d, s_nautilusText);
		}
		break;

	case 7: // close-up spike ball, demo.cpp:762-822
		{
			Spikey_Draw(pDest, timer, delta, true);

			const int dirt = Rocket::geti(trackDirt);
			if (1 != dirt)
				Full(CKD_MULSRC32, d, s_spikeyVignette);

			if (1 == dirt)
			{
				const float raker = Rocket::getf(trackCloseUpMoonraker);
				const float rakerText = ckdh::clampf(0.f, 2.f, Rocket::getf(trackCloseUpMoonrakerText));
				if (raker > 0.f)
				{
					Full(CKD_MULSRC32, d, s_closeSpikeVignetteForRaker);
					Full(CKD_SOFTLIGHT32AA, d, s_closeSpikeDirtRaker, raker);

					if (rakerText > 0.f && rakerText < 1.f)
					{
						CKD_DIRECT(ckd_memset32(s_c, rt2, 0, kOutputSize), "Demo_Draw: memset32");
						Blit(CKD_BLITSRC32, rt2 + (kResY-115)*kResX, s_closeSpike1961.d, kResX, 624, 115);
						Blend(CKD_SOFTLIGHT32AA, d, rt2, kOutputSize, rakerText);
					}
					else if (rakerText >= 1.f)
					{
						const uint32_t *pText = s_closeSpike1961.d;
						const float rakerBlur = ckdh::clampf(0.f, 100.f, Rocket::getf(trackCloseUpMoonrakerTextBlur));
						if (rakerBlur >= 1.f)
						{
							CKD_DIRECT(ckd_old_blur_h(s_c, rt3, pText, 624, 115, ckdh::BoxBlurScale(rakerBlur)), "Demo_Draw: HorizontalBoxBlur32");
							pText = rt3;
						}
						Blit(CKD_BLITSRC32, d + (kResY-115)*kResX, pText, kResX, 624, 115);
					}

					FadeFlash(d, 0.f, fadeToWhite);
					Full(CKD_OVERLAY32, d, s_closeSpikeDirtRaker);
					FadeFlash(d, fadeToBlack, 0.f);
				}
			}
			else if (2 == dirt)
				Full(CKD_SOFTLIGHT32AA, d, s_greetingsDirt, 0.09f*ckdh::kGoldenAngle);
			else if (3 == dirt)
				Full(CKD_SOFTLIGHT32AA, d, s_greetingsDirt, 0.075f*ckdh::kGoldenAngle);

			if (1 != dirt)
				FadeFlash(d, fadeToBlack, fadeToWhite);
		}
		break;

	case 8: // spike ball with title, demo.cpp:824-841
		{
			const int logoIdx = ckdh::clampi(0, 4, Rocket::geti(trackSpikeDemoLogoIndex));
			Spikey_Draw(pDest, timer, delta, false);
			FadeFlash(d, fadeToBlack, fadeToWhite);
			Full(CKD_SOFTLIGHT32, d, s_spikeyBypass);
			Full(CKD_SUB32, d, s_spikeyVignette2);
			Full(CKD_EXCL32, d, s_spikeyFullDirt);
			Full(CKD_MULSRC32A, d, s_vignette06);
			if (0 != logoIdx)
				Full(CKD_MIXOVER32, d, s_spikeyArrested[logoIdx-1]);
			Full(CKD_OVERLAY32, d, s_spikeyVignette);
		}
		break;

	case 9: // free-directional tunnel, demo.cpp:843-852
		{
			Tunnel_Draw(pDest, timer, delta);
			Full(CKD_SUB32, d, s_tunnelVignette2);
			const float show2006 = ckdh::clampf(0.f, 3.f, Rocket::getf(trackShow2006));
			if (show2006 > 0.f)
				Blend(CKD_MIXOVER32, d, LogoBlend(show2006, s_mfx, 4, kResX, kResY), kOutputSize);
		}
		break;

	case 10: // the 'under water' tunnel, demo.cpp:854-874
		{
			const float overlayA = ckdh::saturatef(Rocket::getf(trackWaterLove));
			Sinuses_Draw(pDest, timer, delta);

			const uint32_t *pWaterOverlay = s_waterPrismOverlay.d;
			const float waterOverlayBlurHorz = ckdh::clampf(0.f, 100.f, Rocket::getf(trackLoveBlurHorz));
			if (0.f != waterOverlayBlurHorz)
			{
				CKD_DIRECT(ckd_old_blur_h(s_c, rt0, pWaterOverlay, kResX, kResY, ckdh::BoxBlurScale(waterOverlayBlurHorz)), "Demo_Draw: HorizontalBoxBlur32");
				pWaterOverlay = rt0;
			}
			Blit(CKD_BLITADD32A, d, pWaterOverlay, kResX, kResX, kResY, overlayA);

			if (0 != Rocket::geti(trackDirt))
				Full(CKD_MULSRC32, d, s_waterDirt);

			FadeFlash(d, fadeToBlack, fadeToWhite);
		}
		break;

	case 11: // greetings, demo.cpp:876-891
		{
			Laura_Draw(pDest, timer, delta);

			const int greetSwitch = Rocket::geti(trackGreetSwitch);
			if (greetSwitch < 0 || greetSwitch > 3)
				SetLastError("Demo_Draw: demo:GreetSwitch outside [0, 3]"); // the reference indexes s_pGreetings[] unchecked
			else
				Full(CKD_DARKEN32_50, d, s_greetings[greetSwitch]);
			Full(CKD_SOFTLIGHT32, d, s_greetingsDirt);

			const unsigned yOffs = ((kResY-243)/2) + 227;
			const unsigned xOffs = 24;
			Blit(CKD_BLITSRC32, d + xOffs + yOffs*kResX, s_xboxLogoTPB.d, kResX, 263, 243);

			Full(CKD_OVERLAY32, d, s_greetingsVignette);
		}
		break;

	case 12: // TPB represent, demo.cpp:893-958
		{
			const bool warpAll = 0 != Rocket::geti(trackFullWarpTPB);
			if (false == warpAll)
			{
				CKD_DIRECT(ckd_memset32(s_c, rt0, 0xffffff, kOutputSize), "Demo_Draw: memset32");
				CKD_DIRECT(ckd_memset32(s_c, d, 0xffffff, kOutputSize), "Demo_Draw: memset32");

				const int ribX = ckdh::clampi(0, int(kResX), Rocket::geti(trackRibbonsTPB));
				CKD_DIRECT(ckd_mix_src_s(s_c, d, s_ribbons.d + ribX, kResX, kResY-1, 2160), "Demo_Draw: MixSrc32S");

				Full(CKD_MIXSRC32, rt0, s_nytrikTPB);

				float blurTPB = Rocket::getf(trackBlurTPB);
				if (0.f != blurTPB)
				{
					blurTPB = ckdh::BoxBlurScale(blurTPB);
					CKD_DIRECT(ckd_old_blur_h(s_c, rt0, rt0, kResX, kResY, blurTPB), "Demo_Draw: HorizontalBoxBlur32");
				}
			}
			else
			{
				Plasma_Draw(pDest, timer, delta);

				CKD_DIRECT(ckd_memset32(s_c, rt0, 0xffffff, kOutputSize), "Demo_Draw: memset32");
				Full(CKD_MIXSRC32, rt0, s_nytrikTPB);

				float blurTPB = Rocket::getf(trackBlurTPB);
				if (0.f != blurTPB)
				{
					blurTPB = ckdh::BoxBlurScale(blurTPB);
					CKD_DIRECT(ckd_old_blur_v(s_c, rt0, rt0, kResX, kResY, blurTPB), "Demo_Draw: VerticalBoxBlur32");
				}
			}

			const float distortTPB = Rocket::getf(trackDistortTPB);
			const float distortStrengthTPB = Rocket::getf(trackDistortStrengthTPB);
			CKD_DIRECT(ckd_tape_warp(s_c, rt1, rt0, kResX, kResY, distortStrengthTPB, distortTPB), "Demo_Draw: TapeWarp32");
			Blend(CKD_MIXOVER32, d, rt1, kOutputSize);

			Full(CKD_MULSRC32, d, s_nautilusVignette);
		}
		break;

	case 13: // disco guys and the GPU joke, demo.cpp:960-998
		{
			CKD_DIRECT(ckd_memset32(s_c, d, 0, kOutputSize), "Demo_Draw: memset32");

			const float discoGuys = ckdh::saturatef(Rocket::getf(trackDiscoGuys));
			const float joke = ckdh::saturatef(Rocket::getf(trackCheapJoke));

			if (discoGuys > 0.f)
			{
				const unsigned xStart = (kResX-(8*128))>>1;
				const unsigned yOffs = ((kResY-128)>>1) + 16;
				for (int iGuy = 0; iGuy < 8; ++iGuy)
				{
					const float appearance = ckdh::saturatef(Rocket::getf(trackDiscoGuysAppearance[iGuy]));
					Blit(CKD_BLITSRC32A, d + xStart + iGuy*128 + yOffs*kResX, s_discoGuys[iGuy].d, kResX, 128, 128, discoGuys*smootherstepf(0.f, 1.f, appearance));

					if (discoGuys < 1.f)
					{
						uint32_t *pStrip = d + yOffs*kResX;
						CKD_DIRECT(ckd_old_blur_h(s_c, pStrip, pStrip, kResX, 128, ckdh::BoxBlurScale((1.f-discoGuys)*ckdh::k2PI*ckdh::kGoldenAngle)), "Demo_Draw: HorizontalBoxBlur32");
					}
				}
				Blit(CKD_BLITADD32A, d + (((kResX-1100)/2)-1) + (yOffs+130)*kResX, s_areWeDone.d, kResX, 1100, 57, discoGuys);
			}
			else if (joke > 0.f)
			{
				CKD_DIRECT(ckd_memset32(s_c, d, 0, kOutputSize), "Demo_Draw: memset32");
				Blit(CKD_BLITSRC32A, d + ((kResX-960)/2) + (((kResY-160)/2)*kResX), s_gpuJoke.d, kResX, 960, 160, joke);
			}
		}
		break;

	default:
		DrawTestPattern(d);
	}

	// post fade/flash, demo.cpp:1004-1020
	switch (effect)
	{
	case 1: case 2: case 3: case 6: case 7: case 8: case 10:
		break; // handled by the part
	default:
		FadeFlash(d, fadeToBlack, fadeToWhite);
	}

	FlushChain();
	ckdhost::EndCompose(pDest);
	return true;
}
