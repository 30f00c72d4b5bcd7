// ckd_host.cpp -- C++ host layer: the reference's entry points (include/ckd_host.h) on top of the C ABI (include/ckd.h).
//
// Contains: the GNU Rocket reader (XML project or binary .track files) with sync_get_val's interpolation
// (3rdparty/rocket-stripped/lib/track.c:9-60, device.c:41-78,309-332), the Rocket:: wrapper (rocket.cpp:26-92),
// the per-effect Create/Draw/Destroy shims that read the same tracks the reference reads, and host-buffer wrappers of
// the 2D post ops.

#include "ckd_host_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------------------------
// error + context
// ---------------------------------------------------------------------------------------------------------------

static std::string s_lastError;
void SetLastError(const std::string &description) { s_lastError = description; }
const std::string &CkdHost_GetLastError() { return s_lastError; }

static ckd_ctx *s_ctx = nullptr;
static double s_timeSec = 0.0;
static std::string s_rocketSource;

struct HostImage { std::vector<uint8_t> pixels; int width, height, bpp; };
static std::map<std::string, HostImage> s_images;

static const double kRowRate = (170.0 / (60.0*(170.0/174.0)))*16.0; // audio.cpp:18

static bool Check(int rc, const char *what)
{
	if (CKD_OK == rc)
		return true;
	SetLastError(std::string(what) + ": " + ckd_last_error());
	return false;
}

// Lanes: the host layer is a single context by construction (the reference is not re-entrant either).  For timeline rendering
// a second context on the same device -- own render targets and stream, a copy of every input -- lets two frames be in flight
// side by side; s_ctx is whichever lane the next X_Draw / Demo_Draw renders with.
constexpr int kMaxLanes = 4;
static ckd_ctx *s_lanes[kMaxLanes] = { nullptr, nullptr, nullptr, nullptr };
static int s_device = 0;

bool CkdHost_Create(int resX, int resY, int device)
{
	if (nullptr != s_lanes[0])
		CkdHost_Destroy();
	s_device = device;
	if (!Check(ckd_create(&s_lanes[0], resX, resY, device), "CkdHost_Create"))
		return false;
	s_ctx = s_lanes[0];
	return true;
}

void CkdHost_Destroy()
{
	for (int i = kMaxLanes - 1; i >= 0; --i)
	{
		if (nullptr != s_lanes[i]) ckd_destroy(s_lanes[i]);
		s_lanes[i] = nullptr;
	}
	s_ctx = nullptr;
}

ckd_ctx *CkdHost_Context() { return s_ctx; }
ckd_ctx *CkdHost_LaneContext(int lane) { return (lane >= 0 && lane < kMaxLanes) ? s_lanes[lane] : nullptr; }

// (re)creates the lanes 1..numLanes-1 as copies of the first one's inputs; call after every X_Create / Demo_Create and setter
bool CkdHost_PrepareLanes(int numLanes)
{
	if (nullptr == s_lanes[0])
	{
		SetLastError("CkdHost_Create() has not been called");
		return false;
	}
	if (numLanes < 1 || numLanes > kMaxLanes)
	{
		SetLastError("CkdHost_PrepareLanes: 1 to 4 lanes");
		return false;
	}
	for (int i = 1; i < numLanes; ++i)
	{
		if (nullptr == s_lanes[i])
		{
			if (!Check(ckd_create(&s_lanes[i], ckd_res_x(s_lanes[0]), ckd_res_y(s_lanes[0]), s_device), "CkdHost_PrepareLanes")
				|| !Check(ckd_own_stream(s_lanes[i]), "CkdHost_PrepareLanes"))
				return false;
		}
		if (!Check(ckd_clone_inputs(s_lanes[i], s_lanes[0]), "CkdHost_PrepareLanes"))
			return false;
	}
	return true;
}

bool CkdHost_SelectLane(int lane)
{
	if (lane < 0 || lane >= kMaxLanes || nullptr == s_lanes[lane])
	{
		SetLastError("CkdHost_SelectLane: no such lane");
		return false;
	}
	s_ctx = s_lanes[lane];
	return true;
}

unsigned long long CkdHost_LaunchCount()
{
	unsigned long long n = 0;
	for (ckd_ctx *lane : s_lanes)
		if (nullptr != lane) n += ckd_launch_count(lane);
	return n;
}
void CkdHost_SetRocketSource(const char *path) { s_rocketSource = path ? path : ""; }
void CkdHost_SetTime(double seconds) { s_timeSec = seconds; }

bool CkdHost_PinFrameBuffer(uint32_t *pDest)
{
	if (nullptr == s_ctx)
	{
		SetLastError("CkdHost_Create() has not been called");
		return false;
	}
	return Check(ckd_pin_host(pDest, size_t(ckd_res_x(s_ctx))*ckd_res_y(s_ctx)*sizeof(uint32_t)), "CkdHost_PinFrameBuffer");
}

void CkdHost_UnpinFrameBuffer(uint32_t *pDest) { Check(ckd_unpin_host(pDest), "CkdHost_UnpinFrameBuffer"); }

void CkdHost_RegisterImage(const char *path, const void *pixels, int width, int height, int bytesPerPixel)
{
	HostImage &img = s_images[path];
	img.width = width; img.height = height; img.bpp = bytesPerPixel;
	const size_t bytes = size_t(width)*height*bytesPerPixel;
	img.pixels.assign(static_cast<const uint8_t *>(pixels), static_cast<const uint8_t *>(pixels) + bytes);
}

// a path nobody registered is decoded from its file (host/ckd_image.cpp), like Image_Load32 / Image_Load8 would
static bool DecodeIntoRegistry(const char *path, int bpp)
{
	HostImage img;
	img.bpp = bpp;
	if (nullptr == s_ctx || !ckdhost::DecodeImageForResolution(path, bpp, ckd_res_x(s_ctx), ckd_res_y(s_ctx), img.pixels, img.width, img.height))
		return false;
	s_images[path] = std::move(img);
	return true;
}

// Image_Load32 / Image_Load8 equivalent: registered (or decoded) pixels -> device image slot
static bool LoadImage(const char *path, ckd_image slot, int bpp)
{
	if (nullptr == s_ctx)
	{
		SetLastError("CkdHost_Create() has not been called");
		return false;
	}
	auto it = s_images.find(path);
	if (it == s_images.end() && !DecodeIntoRegistry(path, bpp))
		return false; // DecodeImageFile has set "Can not load image: <path>" (image.cpp:40)
	it = s_images.find(path);
	if (it == s_images.end() || it->second.bpp != bpp)
	{
		SetLastError(std::string("Can not load image: ") + path);
		return false;
	}
	return Check(ckd_set_image(s_ctx, slot, it->second.pixels.data(), it->second.width, it->second.height, bpp), path);
}

// ---------------------------------------------------------------------------------------------------------------
// GNU Rocket
// ---------------------------------------------------------------------------------------------------------------

struct TrackKey { int row; float value; int type; }; // track.h:16-20
struct ckd_sync_track { std::string name; std::vector<TrackKey> keys; };

static std::map<std::string, std::unique_ptr<ckd_sync_track>> s_tracks;     // tracks handed out by AddTrack
static std::map<std::string, std::vector<TrackKey>> s_xmlTracks;           // parsed XML project
static bool s_launched = false;
static double s_rocketRow = 0.0;
static SyncTrack s_stopTrack = nullptr;

// path_encode + sync_track_path, device.c:41-78
static std::string TrackPath(const std::string &base, const std::string &name)
{
	std::string out = base + "_";
	for (unsigned char ch : name)
	{
		if (ch == '.' || ch == '_' || ch == '/' || isalnum(ch))
			out += char(ch);
		else
		{
			out += '-';
			out += "0123456789ABCDEF"[(ch >> 4) & 0xF];
			out += "0123456789ABCDEF"[ch & 0xF];
		}
	}
	return out + ".track";
}

// read_track_data, device.c:309-332
static bool ReadTrackFile(const std::string &path, std::vector<TrackKey> &keys)
{
	FILE *fp = fopen(path.c_str(), "rb");
	if (!fp)
		return false;
	int numKeys = 0;
	if (1 != fread(&numKeys, sizeof(int), 1, fp) || numKeys < 0) { fclose(fp); return false; }
	keys.resize(size_t(numKeys));
	for (auto &key : keys)
	{
		char type = 0;
		if (1 != fread(&key.row, sizeof(int), 1, fp) || 1 != fread(&key.value, sizeof(float), 1, fp) || 1 != fread(&type, sizeof(char), 1, fp))
		{ fclose(fp); return false; }
		key.type = type;
	}
	fclose(fp);
	return true;
}

static bool Attr(const std::string &tag, const char *name, std::string &value)
{
	const std::string needle = std::string(name) + "=\"";
	const size_t at = tag.find(needle);
	if (at == std::string::npos)
		return false;
	const size_t from = at + needle.size();
	const size_t to = tag.find('"', from);
	if (to == std::string::npos)
		return false;
	value = tag.substr(from, to - from);
	return true;
}

// <sync rows=".."><tracks><track name=".."><key interpolation=".." row=".." value=".."/>...
static bool ParseRocketXml(const std::string &path)
{
	FILE *fp = fopen(path.c_str(), "rb");
	if (!fp)
		return false;
	std::string text;
	char buf[65536];
	size_t n;
	while ((n = fread(buf, 1, sizeof(buf), fp)) > 0)
		text.append(buf, n);
	fclose(fp);

	s_xmlTracks.clear();
	std::vector<TrackKey> *current = nullptr;
	size_t pos = 0;
	while ((pos = text.find('<', pos)) != std::string::npos)
	{
		const size_t end = text.find('>', pos);
		if (end == std::string::npos)
			break;
		const std::string tag = text.substr(pos + 1, end - pos - 1);
		pos = end + 1;
		if (0 == tag.compare(0, 6, "track "))
		{
			std::string name;
			if (!Attr(tag, "name", name))
				return false;
			current = &s_xmlTracks[name];
		}
		else if (0 == tag.compare(0, 4, "key ") && current)
		{
			std::string row, value, interp;
			if (!Attr(tag, "row", row) || !Attr(tag, "value", value) || !Attr(tag, "interpolation", interp))
				return false;
			TrackKey key;
			key.row = atoi(row.c_str());
			key.value = strtof(value.c_str(), nullptr);
			key.type = atoi(interp.c_str());
			current->push_back(key);
		}
		else if (0 == tag.compare(0, 6, "/track"))
			current = nullptr;
	}
	return !s_xmlTracks.empty();
}

static bool IsXmlSource() { return s_rocketSource.size() > 7 && 0 == s_rocketSource.compare(s_rocketSource.size() - 7, 7, ".rocket"); }

// sync_find_key + key_idx_floor, track.c:62-85, track.h:29-35
static int KeyIdxFloor(const std::vector<TrackKey> &keys, int row)
{
	int lo = 0, hi = int(keys.size());
	while (lo < hi)
	{
		const int mi = (lo + hi)/2;
		if (keys[mi].row < row) lo = mi + 1;
		else if (keys[mi].row > row) hi = mi;
		else return mi;
	}
	return lo - 1; // -(lo) - 1 negated and biased, then idx = -idx - 2
}

// sync_get_val, track.c:32-60 (double arithmetic, no -ffast-math)
static double SyncGetVal(const ckd_sync_track *t, double row)
{
	const std::vector<TrackKey> &k = t->keys;
	if (k.empty())
		return 0.0;
	const int irow = int(floor(row));
	const int idx = KeyIdxFloor(k, irow);
	if (idx < 0)
		return k[0].value;
	if (idx > int(k.size()) - 2)
		return k[k.size() - 1].value;

	const TrackKey &k0 = k[idx], &k1 = k[idx+1];
	double u = (row - k0.row) / (k1.row - k0.row);
	switch (k0.type)
	{
	case 0: return k0.value;                              // KEY_STEP
	case 1: break;                                        // KEY_LINEAR
	case 2: u = u*u*(3 - 2*u); break;                     // KEY_SMOOTH
	case 3: u = pow(u, 2.0); break;                       // KEY_RAMP
	default: return 0.0;
	}
	return k0.value + (k1.value - k0.value)*u;
}

namespace Rocket
{
	bool Launch()
	{
		s_tracks.clear();
		if (s_rocketSource.empty())
			s_rocketSource = "sync/"; // rocket.cpp:30
		if (IsXmlSource() && !ParseRocketXml(s_rocketSource))
		{
			SetLastError("Can not read GNU Rocket project: " + s_rocketSource);
			return false;
		}
		s_launched = true;
		s_stopTrack = AddTrack("demo:quit"); // rocket.cpp:41
		return true;
	}

	void Land()
	{
		s_tracks.clear();
		s_xmlTracks.clear();
		s_launched = false;
	}

	bool Boost()
	{
		s_rocketRow = s_timeSec*kRowRate; // Audio_Rocket_Sync, audio.cpp:175-178
		if (nullptr != s_stopTrack && 0.0 != get(s_stopTrack))
			return false;
		return true;
	}

	// sync_get_track, device.c:589-611: a track that has no data yet is created empty (value 0)
	SyncTrack AddTrack(const char *name)
	{
		auto it = s_tracks.find(name);
		if (it != s_tracks.end())
			return it->second.get();
		std::unique_ptr<ckd_sync_track> track(new ckd_sync_track);
		track->name = name;
		if (IsXmlSource())
		{
			auto found = s_xmlTracks.find(name);
			if (found != s_xmlTracks.end())
				track->keys = found->second;
		}
		else
		{
			std::string base = s_rocketSource;
			if (!base.empty() && base.back() != '/')
				base += '/';
			ReadTrackFile(TrackPath(base, name), track->keys);
		}
		SyncTrack result = track.get();
		s_tracks[name] = std::move(track);
		return result;
	}

	double get(SyncTrack track) { return SyncGetVal(track, s_rocketRow); }
	int geti(SyncTrack track) { return int(roundf(getf(track))); } // rocket.h:27-29
}

// ---------------------------------------------------------------------------------------------------------------
// frame helpers
// ---------------------------------------------------------------------------------------------------------------

static bool s_pipelined = false;
static int s_slot = 0;

void CkdHost_Flush()
{
	if (nullptr == s_ctx) return;
	Check(ckd_wait_download(s_ctx, 0), "CkdHost_Flush");
	Check(ckd_wait_download(s_ctx, 1), "CkdHost_Flush");
}

void CkdHost_SetPipelined(bool enabled)
{
	if (s_pipelined && !enabled)
		CkdHost_Flush();
	s_pipelined = enabled;
}

static uint32_t *s_composeTarget = nullptr;     // non-null while Demo_Draw composes a frame on the device
static uint32_t *s_deviceTarget = nullptr;      // CkdHost_SetDeviceTarget: where a draw with pDest == nullptr leaves its frame

void CkdHost_SetDeviceTarget(uint32_t *d_frame) { s_deviceTarget = d_frame; }
static int s_readbackBands = -1;                // -1: automatic (4 bands for frames of 8 MB and more), 0/1: off, n: n bands

void CkdHost_SetReadbackBands(int bands) { s_readbackBands = bands; }

// A synchronous X_Draw into a page-locked pDest: let the effect stream its frame to the host in row bands while it is still
// rendering (ckd_arm_readback; the raymarchers without a post chain and the casters that end in a polar remap do, the
// others ignore it).
static void ArmReadback(uint32_t *pDest)
{
	if (nullptr == s_ctx || nullptr == pDest || nullptr != s_composeTarget || s_pipelined)
		return;
	const size_t bytes = size_t(ckd_res_x(s_ctx))*ckd_res_y(s_ctx)*sizeof(uint32_t);
	static const int autoBands = getenv("CKD_READBACK_BANDS") ? atoi(getenv("CKD_READBACK_BANDS")) : 4;
	const int bands = (s_readbackBands >= 0) ? s_readbackBands : (bytes >= (size_t(8) << 20) ? autoBands : 0);
	if (bands >= 2)
		Check(ckd_arm_readback(s_ctx, pDest, bands), "X_Draw");
}

// device buffer the next X_Draw renders into: the single frame twin, or the free slot of the two-deep pipeline
static uint32_t *Target(uint32_t *pDestToStreamTo = nullptr)
{
	ArmReadback(pDestToStreamTo);
	if (nullptr != s_composeTarget)
		return s_composeTarget;
	if (nullptr != s_deviceTarget && nullptr == pDestToStreamTo)
		return s_deviceTarget;
	if (!s_pipelined)
		return ckd_frame(s_ctx);
	Check(ckd_wait_download(s_ctx, s_slot), "X_Draw"); // the copy that last read this buffer must have finished
	return ckd_frame_slot(s_ctx, s_slot);
}

static void Finish(int rc, uint32_t *pDest, const char *what)
{
	int streamed = 0;
	const bool finished = (nullptr == s_ctx) || Check(ckd_finish_readback(s_ctx, &streamed), what); // always: it also disarms
	if (!Check(rc, what) || nullptr != s_composeTarget)
		return;
	if (nullptr == pDest)
		return; // extension: a null pDest leaves the finished frame on the device (ckd_frame / ckd_frame_slot)
	if (finished && streamed)
		return; // the frame arrived in row bands while it was being rendered
	const size_t bytes = size_t(ckd_res_x(s_ctx))*ckd_res_y(s_ctx)*sizeof(uint32_t);
	if (s_pipelined)
	{
		Check(ckd_download_overlapped(s_ctx, pDest, ckd_frame_slot(s_ctx, s_slot), bytes, s_slot), what);
		s_slot ^= 1;
		return;
	}
	if (Check(ckd_download(s_ctx, pDest, ckd_frame(s_ctx), bytes), what))
		Check(ckd_sync(s_ctx), what);
}

namespace ckdhost
{
	bool Check(int rc, const char *what) { return ::Check(rc, what); }

	bool FindImage(const char *path, ImageView &view)
	{
		auto it = s_images.find(path);
		if (it == s_images.end() && DecodeIntoRegistry(path, 4)) // the compositor's art is all Image_Load32 (demo.cpp:198-374)
			it = s_images.find(path);
		if (it == s_images.end())
			return false;
		view = { it->second.pixels.data(), it->second.width, it->second.height, it->second.bpp };
		return true;
	}

	void ReleaseImage(const char *path) { s_images.erase(path); }

	uint32_t *BeginCompose()
	{
		s_composeTarget = nullptr;
		s_composeTarget = Target();
		return s_composeTarget;
	}

	void EndCompose(uint32_t *pDest)
	{
		s_composeTarget = nullptr;
		Finish(CKD_OK, pDest, "Demo_Draw");
	}
}

// ---------------------------------------------------------------------------------------------------------------
// shadertoy.cpp
// ---------------------------------------------------------------------------------------------------------------

static SyncTrack trackLauraSpeed, trackLauraYaw, trackLauraPitch, trackLauraRoll, trackLauraHue, trackLauraSaturate;
static SyncTrack trackNautilusRoll, trackNautilusBlur, trackNautilusHue, trackNautilusSpeed, trackNautilusDesaturation;
static SyncTrack trackSpikeSpeed, trackSpikeRoll, trackSpikeSpecular, trackSpikeDesaturation, trackDistSpikeWarmup, trackSpikeHue, trackSpikeGamma;
static SyncTrack trackDistSpikeX, trackDistSpikeY, trackDistSpikeZ, trackCloseSpikeX, trackCloseSpikeY, trackCloseSpikeZ, trackCloseSpikeZScale;
static SyncTrack trackCloseSpikeNormalGrain, trackCloseSpikeScale, trackCloseSpikeRim, trackCloseSpikeAspectMul;
static SyncTrack trackCloseMixBlurMap, trackCloseMixBlur, trackCloseMixBlurOpacity, trackCloseMixMapBlur;
static SyncTrack trackSinusesSpecular, trackSinusesRoll, trackSinusesSpeed, trackSinusesOffsX, trackSinusesGamma, trackSinusesHue, trackSinusesDesat;
static SyncTrack trackPlasmaSpeed, trackPlasmaHue, trackPlasmaGamma, trackPlasmaDesat;
static SyncTrack trackTunnelBoxy, trackTunnelFlowerScale, trackTunnelFlowerFreq, trackTunnelFlowerPhase, trackTunnelSpeed, trackTunnelRoll, trackTunnelPitch;
static SyncTrack trackTunnelRadius, trackTunnelMulU, trackTunnelMulV, trackTunnelLitTiles, trackTunnelLitBlur, trackTunnelFog1, trackTunnelFog2;

// shadertoy.cpp:101-188
bool Shadertoy_Create()
{
	trackLauraSpeed = Rocket::AddTrack("laura:Speed");
	trackLauraYaw = Rocket::AddTrack("laura:Yaw");
	trackLauraPitch = Rocket::AddTrack("laura:Pitch");
	trackLauraRoll = Rocket::AddTrack("laura:Roll");
	trackLauraHue = Rocket::AddTrack("laura:Hue");
	trackLauraSaturate = Rocket::AddTrack("laura:Saturate");

	trackNautilusRoll = Rocket::AddTrack("nautilus:Roll");
	trackNautilusBlur = Rocket::AddTrack("nautilus:Blur");
	trackNautilusHue = Rocket::AddTrack("nautilus:Hue");
	trackNautilusSpeed = Rocket::AddTrack("nautilus:Speed");
	trackNautilusDesaturation = Rocket::AddTrack("nautilus:Desat");

	trackSpikeSpeed = Rocket::AddTrack("spike:Speed");
	trackSpikeRoll = Rocket::AddTrack("spike:Roll");
	trackSpikeSpecular = Rocket::AddTrack("spike:Specular");
	trackSpikeDesaturation = Rocket::AddTrack("spike:Desaturation");
	trackDistSpikeWarmup = Rocket::AddTrack("distSpike:Warmup");
	trackSpikeHue = Rocket::AddTrack("spike:Hue");
	trackSpikeGamma = Rocket::AddTrack("spike:Gamma");
	trackDistSpikeX = Rocket::AddTrack("distSpike:xOffs");
	trackDistSpikeY = Rocket::AddTrack("distSpike:yOffs");
	trackDistSpikeZ = Rocket::AddTrack("distSpike:zOffs");
	trackCloseSpikeX = Rocket::AddTrack("closeSpike:xOffs");
	trackCloseSpikeY = Rocket::AddTrack("closeSpike:yOffs");
	trackCloseSpikeZ = Rocket::AddTrack("closeSpike:zOffs");
	trackCloseSpikeZScale = Rocket::AddTrack("closeSpike:zOffsScale");
	trackCloseSpikeNormalGrain = Rocket::AddTrack("closeSpike:NormalGrain");
	trackCloseSpikeScale = Rocket::AddTrack("closeSpike:Scale");
	trackCloseSpikeRim = Rocket::AddTrack("closeSpike:Rim");
	trackCloseSpikeAspectMul = Rocket::AddTrack("closeSpike:AspectMul");
	trackCloseMixBlurMap = Rocket::AddTrack("closeSpike:MixBlurMap");
	trackCloseMixBlur = Rocket::AddTrack("closeSpike:MixBlur");
	trackCloseMixMapBlur = Rocket::AddTrack("closeSpike:MixMapBlur");
	trackCloseMixBlurOpacity = Rocket::AddTrack("closeSpike:MixBlurOpacity");

	trackSinusesSpecular = Rocket::AddTrack("sinusesTunnel:Specular");
	trackSinusesRoll = Rocket::AddTrack("sinusesTunnel:Roll");
	trackSinusesSpeed = Rocket::AddTrack("sinusesTunnel:Speed");
	trackSinusesOffsX = Rocket::AddTrack("sinusesTunnel:OffsX");
	trackSinusesGamma = Rocket::AddTrack("sinusesTunnel:Gamma");
	trackSinusesHue = Rocket::AddTrack("sinusesTunnel:Hue");
	trackSinusesDesat = Rocket::AddTrack("sinusesTunnel:Desaturation");

	trackPlasmaSpeed = Rocket::AddTrack("plasma:Speed");
	trackPlasmaHue = Rocket::AddTrack("plasma:Hue");
	trackPlasmaGamma = Rocket::AddTrack("plasma:Gamma");
	trackPlasmaDesat = Rocket::AddTrack("plasma:Desaturation");

	trackTunnelBoxy = Rocket::AddTrack("tunnel:Boxy");
	trackTunnelFlowerScale = Rocket::AddTrack("tunnel:FlowerScale");
	trackTunnelFlowerFreq = Rocket::AddTrack("tunnel:FlowerFreq");
	trackTunnelFlowerPhase = Rocket::AddTrack("tunnel:FlowerPhase");
	trackTunnelSpeed = Rocket::AddTrack("tunnel:Speed");
	trackTunnelRoll = Rocket::AddTrack("tunnel:Roll");
	trackTunnelPitch = Rocket::AddTrack("tunnel:Pitch");
	trackTunnelRadius = Rocket::AddTrack("tunnel:Radius");
	trackTunnelMulU = Rocket::AddTrack("tunnel:MulU");
	trackTunnelMulV = Rocket::AddTrack("tunnel:MulV");
	trackTunnelLitTiles = Rocket::AddTrack("tunnel:LitTiles");
	trackTunnelLitBlur = Rocket::AddTrack("tunnel:LitBlur");
	trackTunnelFog1 = Rocket::AddTrack("tunnel:Fog1");
	trackTunnelFog2 = Rocket::AddTrack("tunnel:Fog2");

	if (!LoadImage("assets/shadertoy/nytrik-hextexture.png", CKD_IMG_TUNNEL_TEX, 4)
		|| !LoadImage("assets/shadertoy/nytrik-hextexture-fx.png", CKD_IMG_TUNNEL_TEX_FX, 4))
		return false;

	// these *must* be FX map sized (shadertoy.cpp:179-183)
	if (!LoadImage("assets/shadertoy/close-up-blur-map-1.png", CKD_IMG_SPIKE_BLUR_MAP0, 4)
		|| !LoadImage("assets/shadertoy/close-up-blur-map-2.png", CKD_IMG_SPIKE_BLUR_MAP1, 4))
		return false;
	for (const char *path : { "assets/shadertoy/close-up-blur-map-1.png", "assets/shadertoy/close-up-blur-map-2.png" })
	{
		const HostImage &img = s_images[path];
		if (img.width != ckd_fxmap_res_x(s_ctx) || img.height != ckd_fxmap_res_y(s_ctx))
		{
			SetLastError(std::string(path) + " must be FX-map sized");
			return false;
		}
	}
	return true;
}

void Shadertoy_Destroy() {}

// shadertoy.cpp:276-280
void Plasma_Draw(uint32_t *pDest, float time, float delta)
{
	(void)delta;
	ckd_plasma_params p;
	p.speed = Rocket::getf(trackPlasmaSpeed);
	p.hue = Rocket::getf(trackPlasmaHue);
	p.gamma = Rocket::getf(trackPlasmaGamma);
	p.desaturation = Rocket::getf(trackPlasmaDesat);
	Finish(ckd_plasma_draw(s_ctx, &p, time, Target(pDest)), pDest, "Plasma_Draw");
}

// shadertoy.cpp:395-407
void Nautilus_Draw(uint32_t *pDest, float time, float delta)
{
	(void)delta;
	ckd_nautilus_params p;
	p.roll = Rocket::getf(trackNautilusRoll);
	p.hue = Rocket::getf(trackNautilusHue);
	p.speed = Rocket::getf(trackNautilusSpeed);
	p.desaturation = Rocket::getf(trackNautilusDesaturation);
	p.blur = Rocket::getf(trackNautilusBlur);
	Finish(ckd_nautilus_draw(s_ctx, &p, time, Target(pDest)), pDest, "Nautilus_Draw");
}

// shadertoy.cpp:661-733
void Spikey_Draw(uint32_t *pDest, float time, float delta, bool close /* = true */)
{
	(void)delta;
	ckd_spikey_params p;
	p.speed = Rocket::getf(trackSpikeSpeed);
	p.roll = Rocket::getf(trackSpikeRoll);
	p.specular = Rocket::getf(trackSpikeSpecular);
	p.desaturation = Rocket::getf(trackSpikeDesaturation);
	p.hue = Rocket::getf(trackSpikeHue);
	p.gamma = Rocket::getf(trackSpikeGamma);
	p.warmup = Rocket::getf(trackDistSpikeWarmup);
	p.dist_x = Rocket::getf(trackDistSpikeX);
	p.dist_y = Rocket::getf(trackDistSpikeY);
	p.dist_z = Rocket::getf(trackDistSpikeZ);
	p.close_x = Rocket::getf(trackCloseSpikeX);
	p.close_y = Rocket::getf(trackCloseSpikeY);
	p.close_z = Rocket::getf(trackCloseSpikeZ);
	p.close_z_scale = Rocket::getf(trackCloseSpikeZScale);
	p.close_normal_grain = Rocket::getf(trackCloseSpikeNormalGrain);
	p.close_scale = Rocket::getf(trackCloseSpikeScale);
	p.close_rim = Rocket::geti(trackCloseSpikeRim);
	p.close_aspect_mul = Rocket::geti(trackCloseSpikeAspectMul);
	p.mix_blur_map = Rocket::getf(trackCloseMixBlurMap);
	p.mix_blur = Rocket::getf(trackCloseMixBlur);
	p.mix_map_blur = Rocket::getf(trackCloseMixMapBlur);
	p.mix_blur_opacity = Rocket::getf(trackCloseMixBlurOpacity);
	Finish(ckd_spikey_draw(s_ctx, &p, time, close ? 1 : 0, Target(pDest)), pDest, "Spikey_Draw");
}

// shadertoy.cpp:840-861
void Tunnel_Draw(uint32_t *pDest, float time, float delta)
{
	(void)delta;
	ckd_tunnel_params p;
	p.boxy = Rocket::getf(trackTunnelBoxy);
	p.flower_scale = Rocket::getf(trackTunnelFlowerScale);
	p.flower_freq = Rocket::getf(trackTunnelFlowerFreq);
	p.flower_phase = Rocket::getf(trackTunnelFlowerPhase);
	p.speed = Rocket::getf(trackTunnelSpeed);
	p.roll = Rocket::getf(trackTunnelRoll);
	p.pitch = Rocket::getf(trackTunnelPitch);
	p.radius = Rocket::getf(trackTunnelRadius);
	p.mul_u = Rocket::getf(trackTunnelMulU);
	p.mul_v = Rocket::getf(trackTunnelMulV);
	p.lit_tiles = Rocket::geti(trackTunnelLitTiles);
	p.lit_blur = Rocket::getf(trackTunnelLitBlur);
	p.fog1 = Rocket::getf(trackTunnelFog1);
	p.fog2 = Rocket::getf(trackTunnelFog2);
	Finish(ckd_tunnel_draw(s_ctx, &p, time, Target()), pDest, "Tunnel_Draw");
}

// shadertoy.cpp:984-988
void Sinuses_Draw(uint32_t *pDest, float time, float delta)
{
	(void)delta;
	ckd_sinuses_params p;
	p.specular = Rocket::getf(trackSinusesSpecular);
	p.roll = Rocket::getf(trackSinusesRoll);
	p.speed = Rocket::getf(trackSinusesSpeed);
	p.offs_x = Rocket::getf(trackSinusesOffsX);
	p.gamma = Rocket::getf(trackSinusesGamma);
	p.hue = Rocket::getf(trackSinusesHue);
	p.desaturation = Rocket::getf(trackSinusesDesat);
	Finish(ckd_sinuses_draw(s_ctx, &p, time, Target(pDest)), pDest, "Sinuses_Draw");
}

// shadertoy.cpp:1105-1109
void Laura_Draw(uint32_t *pDest, float time, float delta)
{
	(void)delta;
	ckd_laura_params p;
	p.speed = Rocket::getf(trackLauraSpeed);
	p.yaw = Rocket::getf(trackLauraYaw);
	p.pitch = Rocket::getf(trackLauraPitch);
	p.roll = Rocket::getf(trackLauraRoll);
	p.hue = Rocket::getf(trackLauraHue);
	p.saturate = Rocket::getf(trackLauraSaturate);
	Finish(ckd_laura_draw(s_ctx, &p, time, Target(pDest)), pDest, "Laura_Draw");
}

// ---------------------------------------------------------------------------------------------------------------
// landscape.cpp
// ---------------------------------------------------------------------------------------------------------------

static SyncTrack trackVoxelScapeForward, trackVoxelScapeTilt, trackWarpSpeed, trackWarpStrength;

// landscape.cpp:194-222
bool Landscape_Create()
{
	if (!LoadImage("assets/scape/D17.png", CKD_IMG_SCAPE_HEIGHT, 1) || !LoadImage("assets/scape/C17W-edit.png", CKD_IMG_SCAPE_COLOR, 4))
		return false;
	if (!LoadImage("assets/scape/foggradient.jpg", CKD_IMG_SCAPE_FOG, 4))
		return false;
	trackVoxelScapeForward = Rocket::AddTrack("voxelScape:Forward");
	trackVoxelScapeTilt = Rocket::AddTrack("voxelScape:Tilt");
	trackWarpSpeed = Rocket::AddTrack("voxelScape:WarpSpeed");
	trackWarpStrength = Rocket::AddTrack("voxelScape:WarpStrength");
	return true;
}

void Landscape_Destroy() {}

// landscape.cpp:228-243; the gamepad (landscape.cpp:112-154) is absent in a headless run: its accumulated state stays zero
void Landscape_Draw(uint32_t *pDest, float time, float delta)
{
	(void)delta;
	ckd_landscape_params p;
	memset(&p, 0, sizeof(p));
	p.forward = Rocket::getf(trackVoxelScapeForward);
	p.tilt = Rocket::getf(trackVoxelScapeTilt);
	p.warp_speed = Rocket::getf(trackWarpSpeed);
	p.warp_strength = Rocket::getf(trackWarpStrength);
	Finish(ckd_landscape_draw(s_ctx, &p, time, Target()), pDest, "Landscape_Draw");
}

// ---------------------------------------------------------------------------------------------------------------
// tunnelscape.cpp
// ---------------------------------------------------------------------------------------------------------------

static SyncTrack trackStarsStepU, trackStarsStepV, trackStarsSpeed, trackStarsBlur;

// tunnelscape.cpp:136-162
bool Tunnelscape_Create()
{
	if (!LoadImage("assets/scape/tscape-D7-edit.png", CKD_IMG_TSCAPE_HEIGHT, 1) || !LoadImage("assets/scape/tscape-C7W-edit.png", CKD_IMG_TSCAPE_COLOR, 4))
		return false;
	if (!LoadImage("assets/scape/foggradient.jpg", CKD_IMG_TSCAPE_FOG, 4))
		return false;
	trackStarsStepU = Rocket::AddTrack("starsTunnel:stepU");
	trackStarsStepV = Rocket::AddTrack("starsTunnel:stepV");
	trackStarsSpeed = Rocket::AddTrack("starsTunnel:Speed");
	trackStarsBlur = Rocket::AddTrack("starsTunnel:Blur");
	return true;
}

void Tunnelscape_Destroy() {}

// tunnelscape.cpp:168-186
void Tunnelscape_Draw(uint32_t *pDest, float time, float delta)
{
	(void)delta;
	ckd_tunnelscape_params p;
	p.step_u = Rocket::getf(trackStarsStepU);
	p.step_v = Rocket::getf(trackStarsStepV);
	p.speed = Rocket::getf(trackStarsSpeed);
	p.blur = Rocket::getf(trackStarsBlur);
	Finish(ckd_tunnelscape_draw(s_ctx, &p, time, Target(pDest)), pDest, "Tunnelscape_Draw");
}

// ---------------------------------------------------------------------------------------------------------------
// ball.cpp
// ---------------------------------------------------------------------------------------------------------------

static SyncTrack trackBallBlur, trackBallRadius, trackBallRayLength, trackBallSpikes, trackBallHasBeams, trackBallBaseShapeIndex, trackBallSpeed;
static SyncTrack trackBallBeamAtten, trackBallBeamAlphaMin, trackBallRotateOffsX, trackBallRotateOffsY, trackBallBeams1, trackBallBeams2, trackBallBeams3, trackBallLowBeams;

// ball.cpp:367-450
static std::vector<uint32_t> s_ballBackground; // s_pBackgrounds[0] for Ball_GetBackground (ball.cpp:412,516)

bool Ball_Create()
{
	static const char *kHeightMapPaths[5] = { "assets/ball/hmap_1_1k.jpg", "assets/ball/hmap_4_1k.jpg", "assets/ball/hmap_2_1k.jpg", "assets/ball/hmap_3_1k.jpg", "assets/ball/hmap_5_1k.jpg" };
	for (int iMap = 0; iMap < 5; ++iMap)
		if (!LoadImage(kHeightMapPaths[iMap], ckd_image(CKD_IMG_BALL_HEIGHT0 + iMap), 1))
			return false;
	if (!LoadImage("assets/ball/colormap_1k.jpg", CKD_IMG_BALL_COLOR0, 4) || !LoadImage("assets/ball/colormap_2_1k.jpg", CKD_IMG_BALL_COLOR1, 4))
		return false;
	if (!LoadImage("assets/ball/beammap_1k_1.jpg", CKD_IMG_BALL_BEAM0, 4) || !LoadImage("assets/ball/beammap_1k_2.jpg", CKD_IMG_BALL_BEAM1, 4)
		|| !LoadImage("assets/ball/beammap_1k_3-2.jpg", CKD_IMG_BALL_BEAM2, 4))
		return false;
	if (!LoadImage("assets/ball/envmap3_1k.jpg", CKD_IMG_BALL_ENV, 4))
		return false;
	if (!LoadImage("assets/ball/nytrik-background_1280x720.png", CKD_IMG_BALL_BACKGROUND0, 4)
		|| !LoadImage("assets/ball/nytrik-background-2-1280x720.png", CKD_IMG_BALL_BACKGROUND1, 4))
		return false;
	if (!LoadImage("assets/ball/halo.png", CKD_IMG_BALL_HALO, 4))
		return false;
	{
		const HostImage &bg = s_images["assets/ball/nytrik-background_1280x720.png"];
		const uint32_t *px = reinterpret_cast<const uint32_t *>(bg.pixels.data());
		s_ballBackground.assign(px, px + size_t(bg.width)*bg.height);
	}

	trackBallBlur = Rocket::AddTrack("ball:Blur");
	trackBallRadius = Rocket::AddTrack("ball:Radius");
	trackBallRayLength = Rocket::AddTrack("ball:RayLength");
	trackBallSpikes = Rocket::AddTrack("ball:Spikes");
	trackBallHasBeams = Rocket::AddTrack("ball:HasBeams");
	trackBallBaseShapeIndex = Rocket::AddTrack("ball:BaseShapeIndex");
	trackBallSpeed = Rocket::AddTrack("ball:Speed");
	trackBallBeamAtten = Rocket::AddTrack("ball:BeamAttenuation");
	trackBallBeamAlphaMin = Rocket::AddTrack("ball:BeamAlphaMin");
	trackBallRotateOffsX = Rocket::AddTrack("ball:RotateOffsX");
	trackBallRotateOffsY = Rocket::AddTrack("ball:RotateOffsY");
	trackBallBeams1 = Rocket::AddTrack("ball:Beams1");
	trackBallBeams2 = Rocket::AddTrack("ball:Beams2");
	trackBallBeams3 = Rocket::AddTrack("ball:Beams3");
	trackBallLowBeams = Rocket::AddTrack("ball:BallLowBeams");
	return true;
}

void Ball_Destroy() { s_ballBackground.clear(); s_ballBackground.shrink_to_fit(); }

// ball.cpp:516-520
uint32_t *Ball_GetBackground() { return s_ballBackground.empty() ? nullptr : s_ballBackground.data(); }

// ball.cpp:452-514
void Ball_Draw(uint32_t *pDest, float time, float delta)
{
	(void)delta;
	ckd_ball_params p;
	p.blur = Rocket::getf(trackBallBlur);
	p.radius = Rocket::getf(trackBallRadius);
	p.ray_length = Rocket::geti(trackBallRayLength);
	p.spikes = Rocket::geti(trackBallSpikes);
	p.has_beams = Rocket::geti(trackBallHasBeams);
	p.base_shape_index = Rocket::geti(trackBallBaseShapeIndex);
	p.speed = Rocket::getf(trackBallSpeed);
	p.beam_atten = Rocket::geti(trackBallBeamAtten);
	p.beam_alpha_min = Rocket::getf(trackBallBeamAlphaMin);
	p.rotate_offs_x = Rocket::getf(trackBallRotateOffsX);
	p.rotate_offs_y = Rocket::getf(trackBallRotateOffsY);
	p.beams1 = Rocket::getf(trackBallBeams1);
	p.beams2 = Rocket::getf(trackBallBeams2);
	p.beams3 = Rocket::getf(trackBallBeams3);
	p.low_beams = Rocket::geti(trackBallLowBeams);
	Finish(ckd_ball_draw(s_ctx, &p, time, Target(pDest)), pDest, "Ball_Draw");
}

bool Ball_HasBeams() { return Rocket::geti(trackBallHasBeams) != 0; } // ball.cpp:521-524

// ---------------------------------------------------------------------------------------------------------------
// torus-twister.cpp
// ---------------------------------------------------------------------------------------------------------------

static SyncTrack trackTwisterSpeed, trackTwisterShearSpeed, trackTwisterBlur;

// torus-twister.cpp:139-160
bool Twister_Create()
{
	if (!LoadImage("assets/twister/hmap_2_1k.jpg", CKD_IMG_TWISTER_HEIGHT, 1) || !LoadImage("assets/twister/colormap_1k.jpg", CKD_IMG_TWISTER_COLOR, 4))
		return false;
	if (!LoadImage("assets/twister/nytrik-background_1280x720.png", CKD_IMG_TWISTER_BACKGROUND, 4))
		return false;
	trackTwisterSpeed = Rocket::AddTrack("twister:Speed");
	trackTwisterShearSpeed = Rocket::AddTrack("twister::ShearSpeed"); // sic (torus-twister.cpp:156)
	trackTwisterBlur = Rocket::AddTrack("twister:Blur");
	return true;
}

void Twister_Destroy() {}

// torus-twister.cpp:166-188
void Twister_Draw(uint32_t *pDest, float time, float delta)
{
	(void)delta;
	ckd_twister_params p;
	p.speed = Rocket::getf(trackTwisterSpeed);
	p.shear_speed = Rocket::getf(trackTwisterShearSpeed);
	p.blur = Rocket::getf(trackTwisterBlur);
	Finish(ckd_twister_draw(s_ctx, &p, time, Target(pDest)), pDest, "Twister_Draw");
}

// ---------------------------------------------------------------------------------------------------------------
// 2D post ops on host buffers: upload -> kernel -> download.  Staging uses the context's render targets 2 and 3,
// which no effect touches.
// ---------------------------------------------------------------------------------------------------------------

namespace {

struct Staged
{
	uint32_t *d_dest = nullptr;
	const uint32_t *d_src = nullptr;
	uint32_t *pDest;
	size_t destBytes;
	bool ok = false;

	Staged(uint32_t *pDest_, size_t destPixels, const uint32_t *pSrc, size_t srcPixels, bool destIsInput)
		: pDest(pDest_), destBytes(destPixels*4)
	{
		if (nullptr == s_ctx) { SetLastError("CkdHost_Create() has not been called"); return; }
		const size_t cap = size_t(ckd_res_x(s_ctx))*ckd_res_y(s_ctx);
		if (destPixels > cap || srcPixels > cap) { SetLastError("buffer larger than the output resolution"); return; }
		d_dest = ckd_render_target(s_ctx, 2);
		if (destIsInput && !Check(ckd_upload(s_ctx, d_dest, pDest, destBytes), "upload")) return;
		if (nullptr == pSrc || (pSrc == pDest && destIsInput))
			d_src = d_dest;
		else
		{
			// also when pSrc == pDest for an op that does not read its destination (TapeWarp32, Polar_Blit, Fx_Blit_2x2 ...):
			// those gather from anywhere in the source, so the source gets its own device copy and the call behaves like the
			// out-of-place one instead of reading whatever render target 2 held
			uint32_t *d = ckd_render_target(s_ctx, 3);
			if (!Check(ckd_upload(s_ctx, d, pSrc, srcPixels*4), "upload")) return;
			d_src = d;
		}
		ok = true;
	}

	void Finish(int rc, const char *what)
	{
		if (ok && Check(rc, what) && Check(ckd_download(s_ctx, pDest, d_dest, destBytes), what))
			Check(ckd_sync(s_ctx), what);
	}
};

void Blend(ckd_blend_op op, uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels, float f, unsigned u, const char *what)
{
	Staged s(pDest, numPixels, pSrc, numPixels, true);
	if (s.ok) s.Finish(ckd_blend(s_ctx, op, s.d_dest, s.d_src, numPixels, f, u), what);
}

size_t OutPixels() { return s_ctx ? size_t(ckd_res_x(s_ctx))*ckd_res_y(s_ctx) : 0; }
size_t FxPixels() { return s_ctx ? size_t(ckd_fxmap_res_x(s_ctx))*ckd_fxmap_res_y(s_ctx) : 0; }

} // namespace

void Polar_Blit(uint32_t *pDest, const uint32_t *pSrc, bool inverse)
{
	Staged s(pDest, OutPixels(), pSrc, OutPixels(), false);
	if (s.ok) s.Finish(ckd_polar_blit(s_ctx, s.d_dest, s.d_src, inverse), "Polar_Blit");
}

void Polar_BlitA(uint32_t *pDest, const uint32_t *pSrc, bool inverse)
{
	Staged s(pDest, OutPixels(), pSrc, OutPixels(), true);
	if (s.ok) s.Finish(ckd_polar_blit_a(s_ctx, s.d_dest, s.d_src, inverse), "Polar_BlitA");
}

void Polar_Blit_2x2(uint32_t *pDest, const uint32_t *pSrc, bool inverse)
{
	Staged s(pDest, FxPixels(), pSrc, FxPixels(), false);
	if (s.ok) s.Finish(ckd_polar_blit_2x2(s_ctx, s.d_dest, s.d_src, inverse), "Polar_Blit_2x2");
}

// fx-blitter.cpp:77-98: stripes into g_pFxMap[0] (the device twin), then Fx_Blit_2x2
void FxBlitter_DrawTestPattern(uint32_t *pDest)
{
	if (nullptr == s_ctx) { SetLastError("CkdHost_Create() has not been called"); return; }
	const unsigned fxX = unsigned(ckd_fxmap_res_x(s_ctx)), fxY = unsigned(ckd_fxmap_res_y(s_ctx));
	std::vector<uint32_t> pattern(size_t(fxX)*fxY);
	for (unsigned iY = 0; iY < fxY; ++iY)
		for (unsigned iX = 0; iX < fxX; ++iX)
			pattern[size_t(iY)*fxX + iX] = (iY < fxY/2) ? ((iY & 1) ? 0xffffffffu : 0u) : ((iX & 1) ? 0xffffffffu : 0u);
	if (nullptr != g_pFxMap[0])
		memcpy(g_pFxMap[0], pattern.data(), pattern.size()*4); // the reference leaves the pattern in g_pFxMap[0]
	uint32_t *d_fx = ckd_fxmap(s_ctx, 0), *d_dest = ckd_render_target(s_ctx, 2);
	if (Check(ckd_upload(s_ctx, d_fx, pattern.data(), pattern.size()*4), "FxBlitter_DrawTestPattern")
		&& Check(ckd_fx_blit_2x2(s_ctx, d_dest, d_fx), "FxBlitter_DrawTestPattern")
		&& Check(ckd_download(s_ctx, pDest, d_dest, OutPixels()*4), "FxBlitter_DrawTestPattern"))
		Check(ckd_sync(s_ctx), "FxBlitter_DrawTestPattern"); // also keeps `pattern` alive until the upload has run
}

void Fx_Blit_2x2(uint32_t *pDest, const uint32_t *pSrc)
{
	Staged s(pDest, OutPixels(), pSrc, FxPixels(), false);
	if (s.ok) s.Finish(ckd_fx_blit_2x2(s_ctx, s.d_dest, s.d_src), "Fx_Blit_2x2");
}

void HorizontalBoxBlur32(uint32_t *pDest, const uint32_t *pSrc, unsigned int xRes, unsigned int yRes, float strength)
{
	Staged s(pDest, size_t(xRes)*yRes, pSrc, size_t(xRes)*yRes, pDest == pSrc);
	if (s.ok) s.Finish(ckd_old_blur_h(s_ctx, s.d_dest, s.d_src, xRes, yRes, strength), "HorizontalBoxBlur32");
}

void VerticalBoxBlur32(uint32_t *pDest, const uint32_t *pSrc, unsigned int xRes, unsigned int yRes, float strength)
{
	Staged s(pDest, size_t(xRes)*yRes, pSrc, size_t(xRes)*yRes, pDest == pSrc);
	if (s.ok) s.Finish(ckd_old_blur_v(s_ctx, s.d_dest, s.d_src, xRes, yRes, strength), "VerticalBoxBlur32");
}

void BoxBlur32(uint32_t *pDest, const uint32_t *pSrc, unsigned int xRes, unsigned int yRes, float strength)
{
	Staged s(pDest, size_t(xRes)*yRes, pSrc, size_t(xRes)*yRes, pDest == pSrc);
	if (s.ok) s.Finish(ckd_old_blur(s_ctx, s.d_dest, s.d_src, xRes, yRes, strength), "BoxBlur32");
}

float BoxBlurScale(float strength) { return ckd_box_blur_scale(strength); }

void BoxBlur_Horz32(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float gain, unsigned numPasses)
{
	Staged s(pDest, size_t(xRes)*yRes, pSrc, size_t(xRes)*yRes, pDest == pSrc);
	if (s.ok) s.Finish(ckd_new_blur_h(s_ctx, s.d_dest, s.d_src, xRes, yRes, strength, gain, numPasses), "BoxBlur_Horz32");
}

void BoxBlur_Vert32(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float gain, unsigned numPasses)
{
	Staged s(pDest, size_t(xRes)*yRes, pSrc, size_t(xRes)*yRes, pDest == pSrc);
	if (s.ok) s.Finish(ckd_new_blur_v(s_ctx, s.d_dest, s.d_src, xRes, yRes, strength, gain, numPasses), "BoxBlur_Vert32");
}

void BoxBlur_32(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float gain, unsigned numPasses)
{
	Staged s(pDest, size_t(xRes)*yRes, pSrc, size_t(xRes)*yRes, pDest == pSrc);
	if (s.ok) s.Finish(ckd_new_blur(s_ctx, s.d_dest, s.d_src, xRes, yRes, strength, gain, numPasses), "BoxBlur_32");
}

void Mix32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels, uint8_t alpha) { Blend(CKD_MIX32, pDest, pSrc, numPixels, 0.f, alpha, "Mix32"); }
void MixOver32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels) { Blend(CKD_MIXOVER32, pDest, pSrc, numPixels, 0.f, 0, "MixOver32"); }
void Add32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels) { Blend(CKD_ADD32, pDest, pSrc, numPixels, 0.f, 0, "Add32"); }
void Sub32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels) { Blend(CKD_SUB32, pDest, pSrc, numPixels, 0.f, 0, "Sub32"); }
void Excl32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels) { Blend(CKD_EXCL32, pDest, pSrc, numPixels, 0.f, 0, "Excl32"); }
void SoftLight32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels) { Blend(CKD_SOFTLIGHT32, pDest, pSrc, numPixels, 0.f, 0, "SoftLight32"); }
void SoftLight32A(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels) { Blend(CKD_SOFTLIGHT32A, pDest, pSrc, numPixels, 0.f, 0, "SoftLight32A"); }
void SoftLight32AA(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels, float alpha) { Blend(CKD_SOFTLIGHT32AA, pDest, pSrc, numPixels, alpha, 0, "SoftLight32AA"); }
void Overlay32(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels) { Blend(CKD_OVERLAY32, pDest, pSrc, numPixels, 0.f, 0, "Overlay32"); }
void Overlay32A(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels) { Blend(CKD_OVERLAY32A, pDest, pSrc, numPixels, 0.f, 0, "Overlay32A"); }
void Darken32_50(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels) { Blend(CKD_DARKEN32_50, pDest, pSrc, numPixels, 0.f, 0, "Darken32_50"); }
void MulSrc32(uint32_t *pDest, const uint32_t *pSrc, unsigned int numPixels) { Blend(CKD_MULSRC32, pDest, pSrc, numPixels, 0.f, 0, "MulSrc32"); }
void MulSrc32A(uint32_t *pDest, const uint32_t *pSrc, unsigned int numPixels) { Blend(CKD_MULSRC32A, pDest, pSrc, numPixels, 0.f, 0, "MulSrc32A"); }
void MixSrc32(uint32_t *pDest, const uint32_t *pSrc, unsigned int numPixels) { Blend(CKD_MIXSRC32, pDest, pSrc, numPixels, 0.f, 0, "MixSrc32"); }
void Fade32(uint32_t *pDest, unsigned int numPixels, uint32_t RGB, uint8_t alpha) { Blend(CKD_FADE32, pDest, nullptr, numPixels, 0.f, (unsigned(alpha) << 24) | (RGB & 0xffffff), "Fade32"); }

// rectangular blits (util.cpp:636-796): pDest usually points INTO a frame (demo.cpp:886 `pDest + xOffs + yOffs*kResX`), so
// the staged extent is the run from the first to the last pixel the op touches; sources may be larger than a frame
// (the 2160-pixel wide ribbons of MixSrc32S, demo.cpp:682) and then get a device buffer of their own.
namespace {

struct StagedRect
{
	uint32_t *d_dest = nullptr, *d_src = nullptr, *d_ownSrc = nullptr;
	uint32_t *pDest;
	size_t destBytes;
	bool ok = false;

	StagedRect(uint32_t *pDest_, size_t destPixels, const uint32_t *pSrc, size_t srcPixels) : pDest(pDest_), destBytes(destPixels*4)
	{
		if (nullptr == s_ctx) { SetLastError("CkdHost_Create() has not been called"); return; }
		if (0 == destPixels || 0 == srcPixels) return;
		if (destPixels > OutPixels()) { SetLastError("buffer larger than the output resolution"); return; }
		d_dest = ckd_render_target(s_ctx, 2);
		if (!Check(ckd_upload(s_ctx, d_dest, pDest, destBytes), "upload")) return;
		if (srcPixels <= OutPixels())
			d_src = ckd_render_target(s_ctx, 3);
		else
		{
			void *p = nullptr;
			if (!Check(ckd_malloc(s_ctx, &p, srcPixels*4), "ckd_malloc")) return;
			d_src = d_ownSrc = static_cast<uint32_t *>(p);
		}
		ok = Check(ckd_upload(s_ctx, d_src, pSrc, srcPixels*4), "upload");
	}

	void Finish(int rc, const char *what)
	{
		if (ok && Check(rc, what) && Check(ckd_download(s_ctx, pDest, d_dest, destBytes), what))
			Check(ckd_sync(s_ctx), what);
	}

	~StagedRect()
	{
		if (d_ownSrc) { ckd_sync(s_ctx); ckd_free(s_ctx, d_ownSrc); }
	}
};

void Blit(ckd_blit_op op, uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned srcResX, unsigned yRes, float alpha, const char *what)
{
	if (0 == yRes || 0 == srcResX) return;
	StagedRect s(pDest, size_t(yRes-1)*destResX + srcResX, pSrc, size_t(srcResX)*yRes);
	if (s.ok) s.Finish(ckd_blit(s_ctx, op, s.d_dest, s.d_src, destResX, srcResX, yRes, alpha), what);
}

} // namespace

void MixSrc32S(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned destResY, unsigned srcStride)
{
	if (0 == destResX || 0 == destResY) return;
	StagedRect s(pDest, size_t(destResX)*destResY, pSrc, size_t(destResY-1)*srcStride + destResX);
	if (s.ok) s.Finish(ckd_mix_src_s(s_ctx, s.d_dest, s.d_src, destResX, destResY, srcStride), "MixSrc32S");
}

void BlitSrc32(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned srcResX, unsigned yRes) { Blit(CKD_BLITSRC32, pDest, pSrc, destResX, srcResX, yRes, 0.f, "BlitSrc32"); }
void BlitSrc32A(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned srcResX, unsigned yRes, float alpha) { Blit(CKD_BLITSRC32A, pDest, pSrc, destResX, srcResX, yRes, alpha, "BlitSrc32A"); }
void BlitAdd32(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned srcResX, unsigned yRes) { Blit(CKD_BLITADD32, pDest, pSrc, destResX, srcResX, yRes, 0.f, "BlitAdd32"); }
void BlitAdd32A(uint32_t *pDest, const uint32_t *pSrc, unsigned destResX, unsigned srcResX, unsigned yRes, float alpha) { Blit(CKD_BLITADD32A, pDest, pSrc, destResX, srcResX, yRes, alpha, "BlitAdd32A"); }

// util.h:57-67
void memset32(void *pDest, int value, size_t numInts)
{
	if (nullptr == s_ctx) { SetLastError("CkdHost_Create() has not been called"); return; }
	uint32_t *d_dest = ckd_render_target(s_ctx, 2);
	for (size_t done = 0; done < numInts; ) // buffers larger than a frame go in frame-sized pieces
	{
		const size_t n = numInts - done < OutPixels() ? numInts - done : OutPixels();
		if (!Check(ckd_memset32(s_ctx, d_dest, uint32_t(value), n), "memset32")
			|| !Check(ckd_download(s_ctx, static_cast<uint32_t *>(pDest) + done, d_dest, n*4), "memset32") || !Check(ckd_sync(s_ctx), "memset32"))
			return;
		done += n;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// module set-up: polar.cpp:61-80, boxblur.cpp:23-36, fx-blitter.cpp:10-25, shared-resources.cpp:14-43
// ---------------------------------------------------------------------------------------------------------------

uint32_t *g_pFxMap[kNumFxMaps] = { nullptr };
uint32_t *g_renderTarget[kNumRenderTargets] = { nullptr };
uint32_t *g_pNytrikTPB = nullptr;
uint32_t *g_pXboxLogoTPB = nullptr;
ckd_unp16 g_gradientUnp16[kNumGradients];

static std::vector<uint32_t> s_nytrikTPB, s_xboxLogoTPB;

static bool HaveContext(const char *what)
{
	if (nullptr != s_ctx) return true;
	SetLastError(std::string(what) + ": CkdHost_Create() has not been called");
	return false;
}

static bool AllocHostImages(uint32_t **images, unsigned count, size_t pixels, const char *what)
{
	for (unsigned i = 0; i < count; ++i)
	{
		if (nullptr != images[i]) continue;
		void *p = nullptr;
		if (!Check(ckd_malloc_host(&p, pixels*sizeof(uint32_t)), what)) return false;
		memset(p, 0, pixels*sizeof(uint32_t));
		images[i] = static_cast<uint32_t *>(p);
	}
	return true;
}

static void FreeHostImages(uint32_t **images, unsigned count)
{
	for (unsigned i = 0; i < count; ++i)
	{
		if (images[i]) ckd_free_host(images[i]);
		images[i] = nullptr;
	}
}

// the polar maps and the blur scratch belong to the context (ckd_create builds s_pMap/s_pInvMap, the 2x2 maps on first use)
bool Polar_Create() { return HaveContext("Polar_Create"); }
void Polar_Destroy() {}
bool BoxBlur_Create() { return HaveContext("BoxBlur_Create"); }
void BoxBlur_Destroy() {}

bool FxBlitter_Create() { return HaveContext("FxBlitter_Create") && AllocHostImages(g_pFxMap, kNumFxMaps, FxPixels(), "FxBlitter_Create"); }
void FxBlitter_Destroy() { FreeHostImages(g_pFxMap, kNumFxMaps); }

static bool CopyImage32(const char *path, std::vector<uint32_t> &out)
{
	auto it = s_images.find(path);
	if (it == s_images.end() && DecodeIntoRegistry(path, 4))
		it = s_images.find(path);
	if (it == s_images.end() || 4 != it->second.bpp)
	{
		SetLastError(std::string("Can not load image: ") + path); // image.cpp:40
		return false;
	}
	const uint32_t *px = reinterpret_cast<const uint32_t *>(it->second.pixels.data());
	out.assign(px, px + size_t(it->second.width)*it->second.height);
	return true;
}

bool Shared_Create()
{
	if (!HaveContext("Shared_Create")) return false;
	for (unsigned i = 0; i < kNumGradients; ++i) // c2vISSE16(i*0x01010101): bytes unpacked into the low four 16-bit lanes
		g_gradientUnp16[i] = { { uint16_t(i), uint16_t(i), uint16_t(i), uint16_t(i), 0, 0, 0, 0 } };
	if (!AllocHostImages(g_renderTarget, kNumRenderTargets, OutPixels(), "Shared_Create")) return false;
	if (!CopyImage32("assets/demo/TPB-logo.png", s_nytrikTPB) || !CopyImage32("assets/demo/tpb_xbox_tp-263x243.png", s_xboxLogoTPB)) return false;
	g_pNytrikTPB = s_nytrikTPB.data();
	g_pXboxLogoTPB = s_xboxLogoTPB.data();
	return true;
}

void Shared_Destroy()
{
	FreeHostImages(g_renderTarget, kNumRenderTargets);
	g_pNytrikTPB = g_pXboxLogoTPB = nullptr;
	s_nytrikTPB.clear(); s_xboxLogoTPB.clear();
}

// fast-cosine.cpp:9-17: the table is the context's (built by ckd_create with the host's cos(), as InitializeFastCosine does)
double g_fastCosTab[kFastCosTabSize+1];

void InitializeFastCosine()
{
	if (nullptr == s_ctx) { SetLastError("CkdHost_Create() has not been called"); return; }
	Check(ckd_get_fast_cos_table(s_ctx, g_fastCosTab), "InitializeFastCosine");
}

static bool FastCosArray(float *pDest, const double *pX, size_t numValues, int sine, const char *what)
{
	if (nullptr == s_ctx) { SetLastError("CkdHost_Create() has not been called"); return false; }
	if (0 == numValues) return true;
	void *d_x = nullptr, *d_out = nullptr;
	bool ok = Check(ckd_malloc(s_ctx, &d_x, numValues*sizeof(double)), what) && Check(ckd_malloc(s_ctx, &d_out, numValues*sizeof(float)), what);
	ok = ok && Check(ckd_upload(s_ctx, d_x, pX, numValues*sizeof(double)), what)
		&& Check(ckd_fastcos(s_ctx, static_cast<float *>(d_out), static_cast<const double *>(d_x), numValues, sine), what)
		&& Check(ckd_download(s_ctx, pDest, d_out, numValues*sizeof(float)), what)
		&& Check(ckd_sync(s_ctx), what);
	if (d_x) ckd_free(s_ctx, d_x);
	if (d_out) ckd_free(s_ctx, d_out);
	return ok;
}

bool fastcosf(float *pDest, const double *pX, size_t numValues) { return FastCosArray(pDest, pX, numValues, 0, "fastcosf"); }
bool fastsinf(float *pDest, const double *pX, size_t numValues) { return FastCosArray(pDest, pX, numValues, 1, "fastsinf"); }

void TapeWarp32(uint32_t *pDest, const uint32_t *pSrc, unsigned xRes, unsigned yRes, float strength, float speed)
{
	Staged s(pDest, size_t(xRes)*yRes, pSrc, size_t(xRes)*yRes, false);
	if (s.ok) s.Finish(ckd_tape_warp(s_ctx, s.d_dest, s.d_src, xRes, yRes, strength, speed), "TapeWarp32");
}

// ---------------------------------------------------------------------------------------------------------------
// extern "C" hooks so ctypes-based tests and bench.py can drive the C++ entry points above
// ---------------------------------------------------------------------------------------------------------------

extern "C" {

int ckdhost_create(int resX, int resY, int device, const char *rocketSource)
{
	if (!CkdHost_Create(resX, resY, device))
		return -1;
	CkdHost_SetRocketSource(rocketSource);
	return 0;
}

// Rocket::Launch + the five X_Create, demo.cpp:140-148
int ckdhost_launch()
{
	if (!Rocket::Launch()) return -2;
	if (!Twister_Create()) return -3;
	if (!Landscape_Create()) return -4;
	if (!Ball_Create()) return -5;
	if (!Tunnelscape_Create()) return -6;
	if (!Shadertoy_Create()) return -7;
	return 0;
}

// forget a registered image, so that the next X_Create decodes the file instead (host/ckd_image.cpp)
void ckdhost_release_image(const char *path) { s_images.erase(path); }

// X_Create of one effect module on its own: 0 twister, 1 landscape, 2 ball, 3 tunnelscape, 4 shadertoy
int ckdhost_effect_create(int module)
{
	s_lastError.clear();
	switch (module)
	{
	case 0: return Twister_Create() ? 0 : -2;
	case 1: return Landscape_Create() ? 0 : -2;
	case 2: return Ball_Create() ? 0 : -2;
	case 3: return Tunnelscape_Create() ? 0 : -2;
	case 4: return Shadertoy_Create() ? 0 : -2;
	}
	return -1;
}

void ckdhost_destroy()
{
	Twister_Destroy(); Landscape_Destroy(); Ball_Destroy(); Tunnelscape_Destroy(); Shadertoy_Destroy();
	Rocket::Land();
	CkdHost_Destroy();
}

// Demo_Create / Demo_Draw / Demo_Destroy, demo.h:8-10
int ckdhost_demo_create() { return Demo_Create() ? 0 : -1; }

int ckdhost_demo_draw(uint32_t *pDest, double seconds, float delta)
{
	s_lastError.clear();
	CkdHost_SetTime(seconds);
	const bool running = Demo_Draw(pDest, float(seconds), delta);
	if (!s_lastError.empty())
		return -2;
	return running ? 1 : 0;
}

void ckdhost_demo_destroy()
{
	Demo_Destroy();
	CkdHost_Destroy();
}

void ckdhost_set_readback_bands(int bands) { CkdHost_SetReadbackBands(bands); }
int ckdhost_pin_frame_buffer(uint32_t *pDest) { return CkdHost_PinFrameBuffer(pDest) ? 0 : -1; }
void ckdhost_unpin_frame_buffer(uint32_t *pDest) { CkdHost_UnpinFrameBuffer(pDest); }

void ckdhost_set_pipelined(int enabled) { CkdHost_SetPipelined(0 != enabled); }
void ckdhost_flush() { CkdHost_Flush(); }
void ckdhost_register_image(const char *path, const void *pixels, int width, int height, int bpp) { CkdHost_RegisterImage(path, pixels, width, height, bpp); }
const char *ckdhost_last_error() { return s_lastError.c_str(); }
void *ckdhost_context() { return s_ctx; }

int ckdhost_set_time(double seconds)
{
	CkdHost_SetTime(seconds);
	return Rocket::Boost() ? 1 : 0;
}

// Rocket only (no GPU needed): used by the CPU-side tests of the track reader
int ckdhost_rocket_open(const char *rocketSource)
{
	CkdHost_SetRocketSource(rocketSource);
	return Rocket::Launch() ? 0 : -1;
}

double ckdhost_track(const char *name) { return Rocket::get(Rocket::AddTrack(name)); }
int ckdhost_track_i(const char *name) { return Rocket::geti(Rocket::AddTrack(name)); }

// same effect ids as oracle/ref_shim.cpp
int ckdhost_draw(int effect, uint32_t *pDest, float time, float delta)
{
	s_lastError.clear();
	switch (effect)
	{
	case 0: Plasma_Draw(pDest, time, delta); break;
	case 1: Nautilus_Draw(pDest, time, delta); break;
	case 2: Spikey_Draw(pDest, time, delta, true); break;
	case 3: Spikey_Draw(pDest, time, delta, false); break;
	case 4: Tunnel_Draw(pDest, time, delta); break;
	case 5: Sinuses_Draw(pDest, time, delta); break;
	case 6: Laura_Draw(pDest, time, delta); break;
	case 7: Landscape_Draw(pDest, time, delta); break;
	case 8: Tunnelscape_Draw(pDest, time, delta); break;
	case 9: Ball_Draw(pDest, time, delta); break;
	case 10: Twister_Draw(pDest, time, delta); break;
	default: return -1;
	}
	return s_lastError.empty() ? 0 : -2;
}

int ckdhost_post(int op, uint32_t *pDest, const uint32_t *pSrc, unsigned a, unsigned b, float f0, float f1, unsigned u)
{
	s_lastError.clear();
	switch (op)
	{
	case 0: Fx_Blit_2x2(pDest, pSrc); break;
	case 1: Polar_Blit(pDest, pSrc, 0 != u); break;
	case 2: Polar_BlitA(pDest, pSrc, 0 != u); break;
	case 3: HorizontalBoxBlur32(pDest, pSrc, a, b, f0); break;
	case 4: VerticalBoxBlur32(pDest, pSrc, a, b, f0); break;
	case 5: BoxBlur32(pDest, pSrc, a, b, f0); break;
	case 6: BoxBlur_32(pDest, pSrc, a, b, f0, f1, u); break;
	case 7: MixSrc32(pDest, pSrc, a); break;
	case 8: SoftLight32(pDest, pSrc, a); break;
	case 9: TapeWarp32(pDest, pSrc, a, b, f0, f1); break;
	case 10: Polar_Blit_2x2(pDest, pSrc, 0 != u); break;
	case 11: FxBlitter_DrawTestPattern(pDest); break;
	case 12: BlitSrc32(pDest, pSrc, a, b, u); break;       // a = destResX, b = srcResX, u = yRes
	case 13: BlitSrc32A(pDest, pSrc, a, b, u, f0); break;
	case 14: BlitAdd32(pDest, pSrc, a, b, u); break;
	case 15: BlitAdd32A(pDest, pSrc, a, b, u, f0); break;
	case 16: MixSrc32S(pDest, pSrc, a, b, u); break;       // a = destResX, b = destResY, u = srcStride
	case 17: memset32(pDest, int(u), size_t(a)); break;
	default: return -1;
	}
	return s_lastError.empty() ? 0 : -2;
}

int ckdhost_fastcos(float *pDest, const double *pX, size_t numValues, int sine)
{
	s_lastError.clear();
	InitializeFastCosine();
	return ((sine ? fastsinf(pDest, pX, numValues) : fastcosf(pDest, pX, numValues)) && s_lastError.empty()) ? 0 : -2;
}

const double *ckdhost_fast_cos_tab() { return g_fastCosTab; }

// module set-up names: 0 Polar, 1 BoxBlur, 2 FxBlitter, 3 Shared; create != 0 calls X_Create, else X_Destroy
int ckdhost_module(int module, int create)
{
	s_lastError.clear();
	bool ok = true;
	switch (module)
	{
	case 0: if (create) ok = Polar_Create(); else Polar_Destroy(); break;
	case 1: if (create) ok = BoxBlur_Create(); else BoxBlur_Destroy(); break;
	case 2: if (create) ok = FxBlitter_Create(); else FxBlitter_Destroy(); break;
	case 3: if (create) ok = Shared_Create(); else Shared_Destroy(); break;
	default: return -1;
	}
	return ok ? 0 : -2;
}

// which: 0-3 g_pFxMap[i], 4-7 g_renderTarget[i], 8 g_pNytrikTPB, 9 g_pXboxLogoTPB, 10 g_gradientUnp16, 11 Ball_GetBackground()
void *ckdhost_global(int which)
{
	if (which >= 0 && which < 4) return g_pFxMap[which];
	if (which >= 4 && which < 8) return g_renderTarget[which-4];
	switch (which)
	{
	case 8: return g_pNytrikTPB;
	case 9: return g_pXboxLogoTPB;
	case 10: return g_gradientUnp16;
	case 11: return Ball_GetBackground();
	}
	return nullptr;
}

} // extern "C"
