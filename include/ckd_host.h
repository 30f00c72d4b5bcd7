// ckd_host.h -- C++ host layer of cookiedough_b200: the reference's own entry points, backed by the CUDA kernels.
//
// Every function below keeps the name, signature, argument meaning and error behaviour of the reference declaration it
// replaces (cited per block, paths relative to the reference's code/ directory), so demo.cpp-style callers compile
// unchanged: X_Create() returns bool and reports through SetLastError(), X_Draw(uint32_t *pDest, float time, float delta)
// fully overwrites the caller-owned HOST buffer pDest (resX*resY ARGB8888), parameters are pulled from the global
// Rocket row exactly where the reference pulls them.  Internally each call fills a POD struct, calls the C ABI
// (include/ckd.h) on device-resident buffers and copies the finished frame back.
//
// The services the reference takes from SDL/BASS/DevIL are replaced by the small CkdHost_* block: resolution and device
// selection (compile-time kResX/kResY in the reference), the time source (BASS stream position), and pre-decoded images.
#pragma once

#include <stdint.h>
#include <string>

#include "ckd.h"

// ---- headless services (replace main.cpp:213-311 start-up, audio.cpp:159-186, image.cpp:31-73) ----------------------

// LUTs, render targets, FX maps, polar maps, blur scratch (main.cpp:263-279) on GPU 'device'; false + SetLastError on failure
bool CkdHost_Create(int resX, int resY, int device);
void CkdHost_Destroy();
ckd_ctx *CkdHost_Context();

// Rocket data source: a GNU Rocket XML project (target/directors-cut.rocket) or a directory with binary "sync/_*.track"
// files (what sync_create_device("sync/") reads, rocket.cpp:30).  Must be set before Rocket::Launch().
void CkdHost_SetRocketSource(const char *path);

// time source of Rocket::Boost(): row = seconds * kRowRate (audio.cpp:18,175-178)
void CkdHost_SetTime(double seconds);

// pre-decoded image for a path the reference would hand to Image_Load32 / Image_Load8 (image.h:10-11);
// bytesPerPixel 4 = BGRA, 1 = luminance.  The pixels are copied.
void CkdHost_RegisterImage(const char *path, const void *pixels, int width, int height, int bytesPerPixel);

// image.h:7-17 -- the reference's loader, without DevIL: PNG and JPEG decoders of the host layer (host/ckd_image.cpp), BGRA
// (0xAARRGGBB) or 8-bit luminance, upper-left origin, pointers collected and freed by Image_Destroy.  X_Create / Demo_Create
// look a path up in the registry first (CkdHost_RegisterImage) and decode the file when it is not there, so a caller that
// runs next to the reference's target/ directory registers nothing.  Relative paths resolve against the asset root
// (default: the working directory, like the reference).
void CkdHost_SetAssetRoot(const char *directory);
bool Image_Create();
void Image_Destroy();
uint32_t *Image_Load32(const std::string &path);
uint8_t *Image_Load8(const std::string &path);
uint32_t *Image_Load32_CA(const std::string &pathC, const std::string &pathA);

// Page-locks the caller's frame buffer in place (pDest of main.cpp:307 is an aligned malloc): X_Draw's copy-back then runs at
// the full PCIe rate instead of through the driver's staging buffers.  Optional; undo before freeing the buffer.
bool CkdHost_PinFrameBuffer(uint32_t *pDest);
void CkdHost_UnpinFrameBuffer(uint32_t *pDest);

// Frame pipelining for offline / timeline rendering (not part of the reference's interface).  While enabled, X_Draw returns
// as soon as the frame is enqueued: pDest is complete after the *second* following X_Draw call or after CkdHost_Flush().
// The device->host copy of frame i then overlaps the rendering of frame i+1 (two device frame buffers, copy stream).
void CkdHost_SetPipelined(bool enabled);

// Synchronous X_Draw into a page-locked pDest (CkdHost_PinFrameBuffer): the raymarched effects without a post chain render
// their frame in row bands and copy every finished band while the next one renders, so most of the rendering hides under
// the PCIe copy.  bands: -1 automatic (4 bands for frames of 8 MB and more; the default), 0 off, n >= 2 that many bands.
void CkdHost_SetReadbackBands(int bands);
void CkdHost_Flush();

// Extension to every X_Draw / Demo_Draw below: pDest == nullptr renders the frame and leaves it on the device
// (ckd_frame(CkdHost_Context()), or the current ckd_frame_slot while pipelined) instead of copying it to the host.

// ... or, after CkdHost_SetDeviceTarget(d_frame), in that device buffer (resX*resY pixels + 4 guard rows; nullptr restores the
// default).  This is how the timeline runner below renders straight into the staging frames of the gather.
void CkdHost_SetDeviceTarget(uint32_t *d_frame);

// Lanes (timeline rendering): up to three more contexts on the same device -- render targets and stream of their own, a copy of
// every input of the first (ckd_clone_inputs) -- so that several frames can be in flight side by side and the latency-bound
// kernels of one (blurs, casters) overlap the issue-bound ones of another.  CkdHost_PrepareLanes(n) (re)creates lanes 1..n-1 from
// the current state of lane 0 (call it after the X_Create / Demo_Create calls); CkdHost_SelectLane picks the context the next
// X_Draw / Demo_Draw uses (lane 0 is the drop-in default).  CkdTimeline_Render does all of this itself when the job asks for lanes.
bool CkdHost_PrepareLanes(int numLanes);
bool CkdHost_SelectLane(int lane);
ckd_ctx *CkdHost_LaneContext(int lane);
unsigned long long CkdHost_LaunchCount();      // kernels launched by both lanes

// ---- timeline rendering, frame-sharded over the GPUs of one box (BASELINE config 5).  One process per GPU calls
//      this with its rank: it renders the frames i with i % world == rank through Demo_Draw -- times[i] is what the reference
//      would get from the audio stream position, audio.cpp:175-178 -- and publishes each to the gather (ckd.h: a slot ring in
//      the collector's HBM, peer copies over NVLink); rank 0 also consumes all frames in order.  Every frame of every pass has
//      the sequence number seqBase + pass*numFrames + i.  The context is put into frame-independent mode for the duration
//      (ckd_set_frame_independent), so the frames -- and their checksums -- do not depend on `world`. -------------------------
struct CkdTimelineJob
{
	const double *times;            // seconds, one per frame
	unsigned numFrames, passes;
	unsigned rank, world;
	ckd_gather *gather;             // nullptr: no exchange, every frame stays on the GPU that rendered it
	int popMode;                    // rank 0: CKD_GATHER_CHECKSUM / CKD_GATHER_TO_HOST bits for ckd_gather_pop
	uint32_t *const *hostRing;      // CKD_GATHER_TO_HOST: page-locked frame buffers used in turn (2..8), or nullptr to
	unsigned hostRingFrames;        //   deliver into the open frame sink (CkdSink_Acquire / CkdSink_Commit by frame index)
	unsigned long long seqBase;
	float delta;                    // Demo_Draw's delta argument
	unsigned lanes;                 // 0 or 1: one frame at a time; 2..4: this rank's frames rotate through that many contexts (see Lanes above)
	unsigned collectorSkip;         // 0 or 1: frame i -> rank i % world.  k > 1: rank 0, which also collects (and checksums / copies
	                                //   out) every frame of every rank, renders only one frame per k rounds of the other ranks:
	                                //   CkdTimeline_Owner(i, world, k).  Every rank must pass the same value.
};
bool CkdTimeline_Render(const CkdTimelineJob *job);
// which rank renders frame i: cycles of k*(world-1) + 1 frames, the first of a cycle goes to rank 0, the others round-robin over
// the ranks 1..world-1 (k <= 1 or world == 1: i % world)
unsigned CkdTimeline_Owner(unsigned frame, unsigned world, unsigned collectorSkip);
// the default for a box: 1 up to 3 GPUs, 2 from 4 GPUs on (measured on B200: at 8 GPUs the collector's own share of the
// rendering plus the per-frame checksum of all eight streams made it the slowest rank)
unsigned CkdTimeline_DefaultCollectorSkip(unsigned world);

// ---- frame sink (replaces Display::Update, display.cpp:66-82, for headless rendering): a raw stream file
//      ("CKDF" header + numFrames ARGB8888 frames) fed through a ring of host buffers by a writer thread.  Acquire a buffer,
//      let X_Draw / Demo_Draw fill it, commit it with its frame index; frames land by index, so several processes (one per
//      GPU) can share one file: one opens it with create = true, the others attach. -----------------------------------------
bool CkdSink_Open(const char *path, unsigned resX, unsigned resY, unsigned numFrames, unsigned ringFrames, bool pinned, bool create);
uint32_t *CkdSink_Acquire();
bool CkdSink_Commit(uint32_t *frame, unsigned frameIndex);
bool CkdSink_Close();

// main.h:49 / main.cpp:175-180
void SetLastError(const std::string &description);
const std::string &CkdHost_GetLastError();

// ---- rocket.h:13-29 ----------------------------------------------------------------------------------------------------

struct ckd_sync_track;
typedef const ckd_sync_track *SyncTrack;

namespace Rocket
{
	bool Launch();
	void Land();
	bool Boost();
	SyncTrack AddTrack(const char *name);
	double get(SyncTrack track);
	inline float getf(SyncTrack track) { return (float) get(track); }
	int geti(SyncTrack track);
}

// ---- shadertoy.h:6-15 -------------------------------------------------------------------------------------------------

bool Shadertoy_Create();
void Shadertoy_Destroy();
void Nautilus_Draw(uint32_t *pDest, float time, float delta);
void Spikey_Draw(uint32_t *pDest, float time, float delta, bool close = true);
void Sinuses_Draw(uint32_t *pDest, float time, float delta);
void Laura_Draw(uint32_t *pDest, float time, float delta);
void Plasma_Draw(uint32_t *pDest, float time, float delta);
void Tunnel_Draw(uint32_t *pDest, float time, float delta);

// ---- landscape.h:7-9, tunnelscape.h:7-9, ball.h:7-13, torus-twister.h:7-9 ---------------------------------------------

bool Landscape_Create();
void Landscape_Destroy();
void Landscape_Draw(uint32_t *pDest, float time, float delta);

bool Tunnelscape_Create();
void Tunnelscape_Destroy();
void Tunnelscape_Draw(uint32_t *pDest, float time, float delta);

bool Ball_Create();
void Ball_Destroy();
void Ball_Draw(uint32_t *pDest, float time, float delta);
uint32_t *Ball_GetBackground(); // ball.cpp:516-520: host pixels of the first background (valid until Ball_Destroy)
bool Ball_HasBeams();

bool Twister_Create();
void Twister_Destroy();
void Twister_Draw(uint32_t *pDest, float time, float delta);

// ---- demo.h:8-10: the compositor.  Demo_Create = Rocket::Launch + the five X_Create + the compositor's tracks and art
//      (every image of demo.cpp:198-374 and shared-resources.cpp:27-34 must have been registered); Demo_Draw advances
//      Rocket itself (CkdHost_SetTime gives the time source), renders the current part's effect and lays the part's art
//      over it on the device, then copies the frame to pDest.  Returns false when the demo is over. -----------------------

bool Demo_Create();
void Demo_Destroy();
bool Demo_Draw(uint32_t *pDest, float time, float delta);

// ---- 2D post chain on HOST buffers (polar.h:7-17, deprecated/boxblur.h:21-41, boxblur.h:7-20, fx-blitter.h:26-33,
//      util.h:57-122).  pDest/pSrc may alias where the reference allows it. ---------------------------------------------

// module set-up of the reference (polar.h:7-8, boxblur.h:7-8, fx-blitter.h:28-29, shared-resources.h:27-28).  The scratch
// images and maps these allocate in the reference live in the device context (CkdHost_Create), so here they check that the
// context exists and provide the caller-visible globals: FxBlitter_Create allocates g_pFxMap[0..3] and Shared_Create
// g_renderTarget[0..3] as page-locked HOST buffers of the reference's sizes (kFxMapBytes / kTargetBytes) for callers that
// use them as scratch between post ops (demo.cpp does); the effects' own intermediate images never leave the device
// (ckd_fxmap / ckd_render_target).  Shared_Create also needs the two TPB logos registered (shared-resources.cpp:27-34).
bool Polar_Create();
void Polar_Destroy();
bool BoxBlur_Create();
void BoxBlur_Destroy();
bool FxBlitter_Create();
void FxBlitter_Destroy();
bool Shared_Create();
void Shared_Destroy();

constexpr unsigned kNumFxMaps = 4;        // fx-blitter.h:14
constexpr unsigned kNumRenderTargets = 4; // shared-resources.h:11
constexpr unsigned kNumGradients = 256;   // shared-resources.h:7
extern uint32_t *g_pFxMap[kNumFxMaps];
extern uint32_t *g_renderTarget[kNumRenderTargets];
extern uint32_t *g_pNytrikTPB;
extern uint32_t *g_pXboxLogoTPB;
// g_gradientUnp16[i] = c2vISSE16(i*0x01010101) (shared-resources.cpp:17-20): eight 16-bit lanes, the low four = i.
// Same 16-byte layout as the reference's __m128i array; filled by Shared_Create.
struct alignas(16) ckd_unp16 { uint16_t lane[8]; };
extern ckd_unp16 g_gradientUnp16[kNumGradients];

// fast (co)sine, fast-cosine.h:9-53 / fast-cosine.cpp:9-17.  g_fastCosTab is filled by InitializeFastCosine() from the
// context's table (ckd_get_fast_cos_table: cos(i*2pi/1024) in double, what the reference computes).  fastcosf / fastsinf over
// an array run on the device with the table staged in shared memory (no effect of the demo calls them -- they exist for PLLs
// in the audio code -- so there is no per-scalar entry: one value per call is not work for a GPU).
constexpr unsigned kFastCosTabSize = 1024;
extern double g_fastCosTab[kFastCosTabSize+1];
void InitializeFastCosine();
bool fastcosf(float *pDest, const double *pX, size_t numValues);
bool fastsinf(float *pDest, const double *pX, size_t numValues);

void Polar_Blit(uint32_t *pDest, const uint32_t *pSrc, bool inverse = false);
void Polar_BlitA(uint32_t *pDest, const uint32_t *pSrc, bool inverse = false);
// FX-map sized buffers (kFxMapResX x kFxMapResY).  The reference's tiles run past the end of the FX map (64 and 32 do not
// divide 644 x 364); this writes exactly kFxMapSize pixels -- what the reference writes in bounds.
void Polar_Blit_2x2(uint32_t *pDest, const uint32_t *pSrc, bool inverse = false);
void Fx_Blit_2x2(uint32_t *pDest, const uint32_t *pSrc);
void FxBlitter_DrawTestPattern(uint32_t *pDest); // fx-blitter.cpp:77-98

void HorizontalBoxBlur32(uint32_t *pDest, const uint32_t *pSrc, unsigned int xRes, unsigned int yRes, float strength);
void VerticalBoxBlur32(uint32_t *pDest, const uint32_t *pSrc, unsigned int xRes, unsigned int yRes, float strength);
void BoxBlur32(uint32_t *pDest, const uint32_t *pSrc, unsigned int xRes, unsigned int yRes, float strength);
float BoxBlurScale(float strength);

void BoxBlur_Horz32(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float gain, unsigned numPasses);
void BoxBlur_Vert32(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float gain, unsigned numPasses);
void BoxBlur_32(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float gain, unsigned numPasses);

void Mix32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels, uint8_t alpha);
void MixOver32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels);
void Add32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels);
void Sub32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels);
void Excl32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels);
void SoftLight32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels);
void SoftLight32A(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels);
void SoftLight32AA(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels, float alpha);
void Overlay32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels);
void Overlay32A(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels);
void Darken32_50(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels);
void MulSrc32(uint32_t *pDest, const uint32_t *pSrc, unsigned int numPixels);
void MulSrc32A(uint32_t *pDest, const uint32_t *pSrc, unsigned int numPixels);
void MixSrc32(uint32_t *pDest, const uint32_t *pSrc, unsigned int numPixels);
void MixSrc32S(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned destResY, unsigned srcStride);
void BlitSrc32(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned srcResX, unsigned yRes);
void BlitSrc32A(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned srcResX, unsigned yRes, float alpha);
void BlitAdd32(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned srcResX, unsigned yRes);
void BlitAdd32A(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned srcResX, unsigned yRes, float alpha);
void Fade32(uint32_t *pDest, unsigned int numPixels, uint32_t RGB, uint8_t alpha);
void memset32(void *pDest, int value, size_t numInts); // util.h:57-67 (numInts a multiple of 4, pDest 8-byte aligned)
void TapeWarp32(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float speed);
