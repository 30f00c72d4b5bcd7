// ckd_raymarch.cu -- the Shadertoy-style raymarchers of shadertoy.cpp as sm_100a kernels.
//
// One thread per FX-map pixel (the reference renders 4 pixels per SSE store; here a pixel is a thread), 32x8 pixel
// tiles walked by a persistent grid (a multiple of the SM count) so the interpolated-cosine LUT is staged into shared
// memory once per CTA.  The march loops are FP32 FMUL/FADD chains (no FMA contraction: the reference has none) plus
// LDS for the LUT; pow/exp/atan2 are the only double-precision ops.  Output goes to the FX map exactly like the
// reference; Fx_Blit_2x2 and the optional blur/blend chain follow as separate HBM-bound kernels (ckd_post.cu).

#include "ckd_internal.h"
#include "ckd_math.cuh"

#include <type_traits>

// unroll factor of the fixed-count march loops (plasma 24, spikey specular-only 36, laura 32 steps)
constexpr int kFixedUnroll = 1; // 2 and 4 measured: no change (the loop overhead is not what limits these kernels)
#include "ckd_hostmath.h"

using namespace ckd;

namespace {

constexpr int kTileX = 32, kTileY = 8;

struct FrameGeom
{
	int fxX, fxY;
	float invFxX, invFxY;     // 1.f/kFxMapResX, 1.f/kFxMapResY (shadertoy-util.h:125-128)
	float aspect, oneOverAspect; // kAspect, kOneOverAspect (main.h:43-44)
};

// Shadertoy::ToUV_FxMap, shadertoy-util.h:123-130
__device__ __forceinline__ void to_uv_fxmap(const FrameGeom &g, unsigned iX, unsigned iY, float scale, float &u, float &v)
{
	float fX = float(iX), fY = float(iY);
	fX *= g.invFxX;
	fY *= g.invFxY;
	u = (fX-0.5f)*scale*g.oneOverAspect;
	v = (fY-0.5f)*scale;
}

struct Rot { float cosine, sine; }; // lutcosf(angle), lutsinf(angle) evaluated once per frame on the host

__device__ __forceinline__ void rot_z(const Rot r, float &X, float &Y) { const float rX = r.cosine*X + r.sine*Y, rY = -r.sine*X + r.cosine*Y; X = rX; Y = rY; }
__device__ __forceinline__ void rot_x(const Rot r, float &Y, float &Z) { const float rY = r.cosine*Y + -r.sine*Z, rZ = r.sine*Y + r.cosine*Z; Y = rY; Z = rZ; }
__device__ __forceinline__ void rot_y(const Rot r, float &X, float &Z) { const float rX = r.cosine*X + r.sine*Z, rZ = -r.sine*X + r.cosine*Z; X = rX; Z = rZ; }

// lerp to a constant per lane: Shadertoy::vLerp4(A, B, f) = A + f*(B-A), shadertoy-util.h:62-67
__device__ __forceinline__ float vlerp(float a, float b, float f) { return a + f*(b-a); }

// Lut = CosLutFast (conversion-free lookups that track their largest angle) or CosLut (exact for every input)
template <class Lut> struct EnvT
{
	Lut lut;             // shared-memory cosine LUT
	RsqrtTab rsqrt;
	FrameGeom geom;
};

// -------------------------------------------------------------------------------------------------------------
// Plasma -- shadertoy.cpp:201-274
// -------------------------------------------------------------------------------------------------------------

struct PlasmaFrame
{
	float time;                 // time*speed
	float dirCos, dirSin;
	float colMulA[3], colMulB[3];
	float gamma;
};

template <class Lut> __device__ __forceinline__ float fPlasma(const Lut &lut, float px, float py, float pz, float time)
{
	const float sine = 0.2f*lutsinf(lut, px-py);
	const float fX = sine + lutcosf(lut, px*0.33f);
	const float fY = sine + lutcosf(lut, py*0.43f);
	const float fZ = sine + lutcosf(lut, (5.f*time+pz)*0.53f);
	return sqrtf(fX*fX + fY*fY + fZ*fZ)-0.8f;
}

struct PlasmaEffect
{
	PlasmaFrame f;
	template <class Lut> __device__ __forceinline__ uint32_t shade(const EnvT<Lut> &e, unsigned iX, unsigned iY) const
	{
		float u, v;
		to_uv_fxmap(e.geom, iX, iY, 4.f, u, v);

		const float dx = f.dirCos*u*e.geom.aspect - f.dirSin*0.75f;
		const float dy = v;
		const float dz = f.dirSin*u + f.dirCos*0.75f;

		float total = 0.f, march = 0.f;
		float hx = 0.f, hy = 0.f, hz = 0.f;
		#pragma unroll kFixedUnroll
		for (int iStep = 0; iStep < 24; ++iStep)
		{
			march = fPlasma(e.lut, hx, hy, hz, f.time);
			total += march*(0.5f*kGoldenRatio);
			hx = dx*total;
			hy = dy*total;
			hz = dz*total;
		}

		const float second = fPlasma(e.lut, hx*0.5f, hy*0.5f, hz*0.5f, f.time);
		const float mul = 8.f - dx*0.5f;
		const float b = (f.colMulA[0]*march + f.colMulB[0]*second)*mul;
		const float g = (f.colMulA[1]*march + f.colMulB[1]*second)*mul;
		const float r = (f.colMulA[2]*march + f.colMulB[2]*second)*mul;

		// the 4th lane of a Vector3 colour is 0: log_ps(0) = NaN -> exp_ps -> huge -> converts to 0
		// (a lane holding 0 -> log_ps = NaN -> exp_ps = e^88 -> x255 = inf -> cvtps2dq = 0x80000000 -> max(0, .) = 0: alpha is 0)
		return to_pixel(gamma_adj1(b, f.gamma), gamma_adj1(g, f.gamma), gamma_adj1(r, f.gamma), 0.f);
	}
};

// -------------------------------------------------------------------------------------------------------------
// Nautilus -- shadertoy.cpp:288-393
// -------------------------------------------------------------------------------------------------------------

struct NautilusFrame
{
	float time;               // time*speed
	float gx, gy, gz;         // fNautilus_global
	float diffColor[3];
	float cosHitOffs, funkCos;
	Rot roll;
};

template <class Lut> __device__ __forceinline__ float fNautilus(const Lut &lut, const NautilusFrame &f, float px, float py, float pz)
{
	const float cosX = lutcosf(lut, lutcosf(lut, px + f.gx)*px - lutcosf(lut, py + f.gy)*py);
	const float cosY = lutcosf(lut, pz*0.33f*px - f.gz*py);
	const float cosZ = lutcosf(lut, px + py + pz*0.8f + f.time);
	const float dotted = cosX*cosX + cosY*cosY + cosZ*cosZ;
	return dotted*0.5f - .7f;
}

struct NautilusEffect
{
	NautilusFrame f;
	template <class Lut> __device__ __forceinline__ uint32_t shade(const EnvT<Lut> &e, unsigned iX, unsigned iY) const
	{
		float u, v;
		to_uv_fxmap(e.geom, iX, iY, 2.f, u, v);

		vec3 dir = { u*e.geom.aspect, v, 1.f };
		rot_z(f.roll, dir.x, dir.y);
		fast_norm3(e.rsqrt, dir);

		float hx = 0.f, hy = 0.f, hz = 0.f;
		float total = 0.01f, march = 1.f;
		#pragma unroll 1
		for (int iStep = 0; march > 0.01f && iStep < 48; ++iStep)
		{
			hx = dir.x*total;
			hy = dir.y*total;
			hz = dir.z*total;
			march = fNautilus(e.lut, f, hx, hy, hz);
			total += march*0.628f;
		}

		constexpr float nOffs = 0.15f;
		vec3 normal = {
			march-fNautilus(e.lut, f, hx+nOffs, hy, hz),
			march-fNautilus(e.lut, f, hx, hy+nOffs, hz),
			march-fNautilus(e.lut, f, hx, hy, hz+nOffs) };
		fast_norm3(e.rsqrt, normal);

		float diffuse = normal.z*0.1f;
		const float specular = powf_ref(stdmax(0.f, dot3(normal, dir)), 16.f);

		const float ox = hx + f.cosHitOffs, oy = hy + f.cosHitOffs, oz = hz + f.cosHitOffs;
		vec3 funk = {
			march-fNautilus(e.lut, f, ox+nOffs, oy, oz),
			march-fNautilus(e.lut, f, ox, oy+nOffs, oz),
			march-fNautilus(e.lut, f, ox, oy, oz+nOffs) };
		fast_norm3(e.rsqrt, funk);

		const float yMod = fracf(hy*0.3f + funk.x*0.628f + funk.y*f.funkCos);
		diffuse *= yMod*yMod*yMod;

		const float s = 1.56f*total + specular;
		const float add = specular*kGoldenRatio*0.2f;
		const float b = (diffuse + f.diffColor[0]*s) + add;
		const float g = (diffuse + f.diffColor[1]*s) + add;
		const float r = (diffuse + f.diffColor[2]*s) + add;

		return to_pixel(gamma_adj1(b, 1.44f), gamma_adj1(g, 1.44f), gamma_adj1(r, 1.44f), 0.f); // Vector3 colour: 4th lane 0 -> alpha 0
	}
};

// -------------------------------------------------------------------------------------------------------------
// Spikey (close / distant / specular only) -- shadertoy.cpp:416-659
// -------------------------------------------------------------------------------------------------------------

struct SpikeyFrame
{
	float gx, gy, gz;         // fSpike_global.xyz
	float diffColor[4];       // all four lanes feed the pixel
	Rot roll;
	float specPow, gamma;
	float xOffs, yOffs, zTerm; // close: dir.z = 1+zOffsFinal; distant: origin.z = -2.614+zOffs
	float normalGrain;
	float warmup;
};

template <bool GOLDEN_ANGLE, class Lut>
__device__ __forceinline__ float fSpikey(const Lut &lut, const SpikeyFrame &f, float px, float py, float pz)
{
	// fSpikey1 (kGoldenAngle) / fSpikey2 (kGoldenRatio), shadertoy.cpp:418-430
	constexpr float scale = (GOLDEN_ANGLE ? kGoldenAngle : kGoldenRatio)*0.1f;
	const float radius = 1.35f + scale*lutcosf(lut, f.gy*py - f.gx) + scale*lutcosf(lut, f.gz*px + f.gx);
	const vec3 p = { px, py, pz };
	return fast_len3(p) - radius;
}

struct SpikeyCloseEffect
{
	SpikeyFrame f;
	template <class Lut> __device__ __forceinline__ uint32_t shade(const EnvT<Lut> &e, unsigned iX, unsigned iY) const
	{
		float u, v;
		to_uv_fxmap(e.geom, iX, iY, 2.f, u, v);

		const float ox = 0.2f, oy = 0.f, oz = -2.23f;
		vec3 dir = { (u+f.xOffs)*e.geom.aspect, v + f.yOffs, f.zTerm };
		rot_z(f.roll, dir.x, dir.y);
		fast_norm3(e.rsqrt, dir);

		float hx = 0.f, hy = 0.f, hz = 0.f;
		float march = 1.f, total = 0.f;
		#pragma unroll 1
		for (int iStep = 0; march > 0.0001f && iStep < 32; ++iStep)
		{
			hx = ox + dir.x*total;
			hy = oy + dir.y*total;
			hz = oz + dir.z*total;
			march = fSpikey<true>(e.lut, f, hx, hy, hz);
			total += march*(0.05f*kPI);
		}

		const float nOffs = f.normalGrain;
		vec3 normal = {
			march-fSpikey<true>(e.lut, f, hx+nOffs, hy, hz),
			march-fSpikey<true>(e.lut, f, hx, hy+nOffs, hz),
			march-fSpikey<true>(e.lut, f, hx, hy, hz+nOffs) };
		fast_norm3(e.rsqrt, normal);

		// the 'rim' branch (shadertoy.cpp:504-511) multiplies by max(1, min(0, rim)) == 1: a no-op kept out of the kernel
		const float diffuse = normal.z;
		const float specular = powf_ref(stdmax(0.f, dot3(normal, dir)), f.specPow);
		const float distance = hz-oz;
		const float fog = exp_fog(distance, kGoldenRatio*0.1f);

		float c[4];
		#pragma unroll
		for (int i = 0; i < 4; ++i)
			c[i] = gamma_adj1(vlerp((f.diffColor[i] + specular)*diffuse, 1.f, fog), f.gamma);
		return to_pixel(c[0], c[1], c[2], c[3]);
	}
};

struct SpikeyDistantEffect
{
	SpikeyFrame f;
	template <class Lut> __device__ __forceinline__ uint32_t shade(const EnvT<Lut> &e, unsigned iX, unsigned iY) const
	{
		float u, v;
		to_uv_fxmap(e.geom, iX, iY, 2.f, u, v);

		const float ox = 0.f, oy = 0.f, oz = f.zTerm;
		vec3 dir = { u + f.xOffs, v + f.yOffs, 1.f };
		rot_z(f.roll, dir.x, dir.y);
		fast_norm3(e.rsqrt, dir);

		float hx = 0.f, hy = 0.f, hz = 0.f;
		float march = 1.f, total = 0.f;
		#pragma unroll 1
		for (int iStep = 0; march > 0.001f && iStep < 48; ++iStep)
		{
			hx = ox + dir.x*total;
			hy = oy + dir.y*total;
			hz = oz + dir.z*total;
			march = fSpikey<false>(e.lut, f, hx, hy, hz);
			march *= 0.314f;
			total += march;
		}

		constexpr float nOffs = kPI*0.02f;
		vec3 normal = {
			march-fSpikey<false>(e.lut, f, hx+nOffs, hy, hz),
			march-fSpikey<false>(e.lut, f, hx, hy+nOffs, hz),
			march-fSpikey<false>(e.lut, f, hx, hy, hz+nOffs) };
		fast_norm3(e.rsqrt, normal);

		const float diffuse = stdmax(0.f, normal.z*0.8f + normal.y*0.2f);
		const float fakeSpecular = powf_ref(dot3(normal, dir), f.specPow); // unclamped base on purpose (shadertoy.cpp:586)
		const float distance = hz-oz;
		const float fog = exp_fog(distance, 0.133f);

		float c[4];
		#pragma unroll
		for (int i = 0; i < 4; ++i)
			c[i] = gamma_adj1(vlerp((f.diffColor[i] + fakeSpecular)*diffuse, 1.f, fog), f.gamma);
		return to_pixel(c[0], c[1], c[2], c[3]);
	}
};

struct SpikeySpecOnlyEffect
{
	SpikeyFrame f;
	template <class Lut> __device__ __forceinline__ uint32_t shade(const EnvT<Lut> &e, unsigned iX, unsigned iY) const
	{
		float u, v;
		to_uv_fxmap(e.geom, iX, iY, kGoldenRatio, u, v);

		const float ox = 0.f, oy = 0.f, oz = -3.314f;
		vec3 dir = { u*e.geom.aspect, v, 1.f };
		rot_z(f.roll, dir.x, dir.y);
		fast_norm3(e.rsqrt, dir);

		float hx = 0.f, hy = 0.f, hz = 0.f;
		float march = 1.f, total = 0.f;
		#pragma unroll kFixedUnroll
		for (int iStep = 0; iStep < 36; ++iStep)
		{
			hx = ox + dir.x*total;
			hy = oy + dir.y*total;
			hz = oz + dir.z*total;
			march = fSpikey<false>(e.lut, f, hx, hy, hz);
			total += march*0.075f*kGoldenRatio;
		}

		constexpr float nOffs = 0.01f;
		vec3 normal = {
			march-fSpikey<false>(e.lut, f, hx+nOffs, hy, hz),
			march-fSpikey<false>(e.lut, f, hx, hy+nOffs, hz),
			march-fSpikey<false>(e.lut, f, hx, hy, hz+nOffs) };
		fast_norm3(e.rsqrt, normal);

		const float distance = hz-oz;
		const float fakeSpecular = f.warmup*powf_ref(stdmax(0.f, dot3(normal, dir)), f.specPow);
		const float fogged = vlerp(fakeSpecular, 0.f, exp_fog(distance, 0.0133f));
		const uint32_t chan = to_chan(fogged);
		return chan*0x01010101u;
	}
};

// -------------------------------------------------------------------------------------------------------------
// Sinuses -- shadertoy.cpp:870-982
// -------------------------------------------------------------------------------------------------------------

struct SinusesFrame
{
	float origin[3];
	float diffColor[4];
	Rot roll;
	float specPow, gamma, offsX;
};

template <class Lut> __device__ __forceinline__ float fSinMap(const Lut &lut, float px0, float py0, float pZ)
{
	const float zMod = pZ*0.314f;
	const float pathCos = lutcosf(lut, zMod);
	const float pathCos2 = lutcosf(lut, zMod+(k2PI/4.f))*kGoldenRatio;
	const float pX = px0-(pathCos2*2.f - pathCos*1.5f);
	const float pY = py0-(pathCos*3.14f + pathCos2);

	const float aX = pX*0.315f*1.25f + lutsinf(lut, pZ*(0.814f*1.25f));
	const float aY = pY*0.315f*1.25f + lutsinf(lut, pX*(0.814f*1.25f));
	const float aZ = pZ*0.315f*1.25f + lutsinf(lut, pY*(0.814f*1.25f));

	const float cosX = lutcosf(lut, aX);
	const float cosY = lutcosf(lut, aY);
	const float cosZ = lutcosf(lut, aZ);

	const float length = sqrtf(cosX*cosX + cosY*cosY + cosZ*cosZ);
	return (length - 1.025f)*1.33f;
}

struct SinusesEffect
{
	SinusesFrame f;
	template <class Lut> __device__ __forceinline__ uint32_t shade(const EnvT<Lut> &e, unsigned iX, unsigned iY) const
	{
		float u, v;
		to_uv_fxmap(e.geom, iX, iY, 2.f, u, v);

		vec3 dir = { (u+f.offsX)*e.geom.aspect, v, 0.314f };
		rot_z(f.roll, dir.x, dir.y);
		fast_norm3(e.rsqrt, dir);

		float hx = 0.f, hy = 0.f, hz = 0.f;
		float march = 1.f, total = 0.f;
		#pragma unroll 1
		for (int iStep = 0; march > 0.01f && iStep < 32; ++iStep)
		{
			hx = f.origin[0] + dir.x*total;
			hy = f.origin[1] + dir.y*total;
			hz = f.origin[2] + dir.z*total;
			march = fSinMap(e.lut, hx, hy, hz);
			total += march*0.814f;
		}

		constexpr float nOffs = 0.2f;
		vec3 normal = {
			march-fSinMap(e.lut, hx+nOffs, hy, hz),
			march-fSinMap(e.lut, hx, hy+nOffs, hz),
			march-fSinMap(e.lut, hx, hy, hz+nOffs) };
		fast_norm3(e.rsqrt, normal);

		float diffuse = normal.z*0.7f + 0.3f*normal.y;
		diffuse = 0.2f + 0.8f*diffuse;

		const float specDot = dot3(normal, dir);
		const float fakeSpecular = powf_ref(specDot, f.specPow);
		const float distance = hz-f.origin[2];
		const float fog = exp_fog(distance, 0.03f);

		float c[4];
		#pragma unroll
		for (int i = 0; i < 4; ++i)
			c[i] = gamma_adj1(vlerp((f.diffColor[i] + fakeSpecular)*diffuse, 1.f, fog), f.gamma);
		return to_pixel(c[0], c[1], c[2], c[3]);
	}
};

// -------------------------------------------------------------------------------------------------------------
// Laura -- shadertoy.cpp:998-1103, Shadertoy::Specular shadertoy-util.h:252-280
// -------------------------------------------------------------------------------------------------------------

struct LauraFrame
{
	float originZ;
	float diffColor[4];
	Rot yaw, pitch, roll;
};

template <class Lut> __device__ __forceinline__ float fLaura(const Lut &lut, float px, float py, float pz)
{
	return lutcosf(lut, px)+lutcosf(lut, py)+lutcosf(lut, pz) + 1.f;
}

struct LauraEffect
{
	LauraFrame f;
	template <class Lut> __device__ __forceinline__ uint32_t shade(const EnvT<Lut> &e, unsigned iX, unsigned iY) const
	{
		float u, v;
		to_uv_fxmap(e.geom, iX, iY, 2.f, u, v);

		vec3 dir = { u*e.geom.aspect, v, kPI };
		rot_y(f.yaw, dir.x, dir.z);
		rot_x(f.pitch, dir.y, dir.z);
		rot_z(f.roll, dir.x, dir.y);
		fast_norm3(e.rsqrt, dir);

		const vec3 origin = { 0.f, 0.f, f.originZ };
		vec3 hit = { 0.f, 0.f, 0.f };
		float march = 0.f, total = 0.f;
		#pragma unroll kFixedUnroll
		for (int iStep = 0; iStep < 32; ++iStep)
		{
			hit.x = origin.x + dir.x*total;
			hit.y = origin.y + dir.y*total;
			hit.z = origin.z + dir.z*total;
			march = fLaura(e.lut, hit.x, hit.y, hit.z);
			total += march*0.5f;
		}

		// LauraNormal, shadertoy.cpp:1003-1015
		constexpr float nOffs = 0.1628f;
		vec3 normal = {
			fLaura(e.lut, hit.x+nOffs, hit.y, hit.z)-march,
			fLaura(e.lut, hit.x, hit.y+nOffs, hit.z)-march,
			fLaura(e.lut, hit.x, hit.y, hit.z+nOffs)-march };
		fast_norm3(e.rsqrt, normal);

		const vec3 lightPos = { origin.x-dir.x, origin.y-dir.y, origin.z-dir.z };
		vec3 lightDir = { lightPos.x-hit.x, lightPos.y-hit.y, lightPos.z-hit.z };
		fast_norm3(e.rsqrt, lightDir);

		float diffuse = stdmax(0.3f, dot3(normal, lightDir));
		const float distance = hit.z-origin.z;

		// Shadertoy::Specular(origin, hit, normal, lightDir, 4.f)
		float specular;
		{
			vec3 V = { origin.x-hit.x, origin.y-hit.y, origin.z-hit.z };
			const float oneOverLenV = rsqrt_x86(e.rsqrt, dp_ps3(V, V));
			V.x *= oneOverLenV; V.y *= oneOverLenV; V.z *= oneOverLenV;
			vec3 H = { lightDir.x+V.x, lightDir.y+V.y, lightDir.z+V.z };
			const float oneOverLenH = rsqrt_x86(e.rsqrt, dp_ps3(H, H));
			H.x *= oneOverLenH; H.y *= oneOverLenH; H.z *= oneOverLenH;
			const float cosAng = dp_ps3(normal, H);
			specular = (0 == (__float_as_uint(cosAng) >> 31)) ? powf_ref(cosAng, 4.f) : 0.f;
		}

		// rim (shadertoy.cpp:1085-1088) evaluates to max(1, min(0, rim)) == 1
		diffuse *= 1.f;

		const float fogColor = q3_rsqrtf2(specular+diffuse);
		const float fog = exp_fog(distance, 0.001f);
		const float lit = diffuse+specular;

		float c[4];
		#pragma unroll
		for (int i = 0; i < 4; ++i)
			c[i] = gamma_adj1(vlerp(f.diffColor[i]*lit, fogColor, fog), 1.44f);
		return to_pixel(c[0], c[1], c[2], c[3]);
	}
};

// -------------------------------------------------------------------------------------------------------------
// generic persistent tile kernel
// -------------------------------------------------------------------------------------------------------------

// Work distribution: the FX map is cut into 8x4-pixel warp tiles (each row of a tile is one 32-byte sector) and every warp
// of a persistent grid pulls the next tile index from a global counter.  Per-tile cost varies (early-exit march loops) and
// a static split leaves ~1/7 of the machine idle in the last round at 4K; pulling 32-pixel units keeps every scheduler busy
// to the end.  Two counters alternate between launches: launch k consumes counter[k&1] and zeroes the other one.
constexpr int kWarpTileX = 8, kWarpTileY = 4;

struct TileQueue { unsigned *counter; unsigned *nextCounter; int tilesX; unsigned firstTile, numTiles; }; // tiles [firstTile, numTiles)

__device__ __forceinline__ bool next_tile(const TileQueue &q, unsigned &iX, unsigned &iY)
{
	// (requesting the next tile's number one tile ahead, so that the atomic's round trip overlaps the shading, was measured
	//  and made every kernel 2-6 % slower -- profiles/r02_notes.md)
	const unsigned lane = threadIdx.x;
	unsigned tile = 0;
	if (lane == 0)
		tile = q.firstTile + atomicAdd(q.counter, 1u);
	tile = __shfl_sync(0xffffffffu, tile, 0);
	if (tile >= q.numTiles)
		return false;
	const unsigned tY = tile / unsigned(q.tilesX), tX = tile - tY*unsigned(q.tilesX);
	iX = tX*kWarpTileX + (lane & 7);
	iY = tY*kWarpTileY + (lane >> 3);
	return true;
}

// FAST = true: the conversion-free LUT lookups (CosLutFast), launched only for frames whose LUT angles the host has proved
// in range (LutRangeProof below); FAST = false: the exact lookups, bit-exact for every input.
template <class Effect, bool FAST>
__global__ void __launch_bounds__(kTileX*kTileY) raymarch_kernel(const Effect effect, uint32_t *__restrict__ pDest, const FrameGeom geom,
	const float2 *__restrict__ g_lut2, const RsqrtTab rsqrt, const TileQueue queue)
{
	__shared__ float2 s_lut2[2048];
	if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0)
		*queue.nextCounter = 0;
	const CosLut lut = stage_cos_lut(s_lut2, g_lut2);

	const typename std::conditional<FAST, EnvT<CosLutFast>, EnvT<CosLut>>::type env = { { lut }, rsqrt, geom };

	unsigned iX, iY;
	while (next_tile(queue, iX, iY))
	{
		if (iX < unsigned(geom.fxX) && iY < unsigned(geom.fxY))
			pDest[size_t(iY)*geom.fxX + iX] = effect.shade(env, iX, iY);
	}
}

// -------------------------------------------------------------------------------------------------------------
// Free-directional tunnel -- shadertoy.cpp:746-838 (writes two maps)
// -------------------------------------------------------------------------------------------------------------

struct TunnelFrame
{
	float boxy, flowerScale, flowerFreq, flowerPhase;
	float timeSpeed;          // (time*speed)*speed: 'time *= speed' then 'time*speed' (shadertoy.cpp:767,802)
	Rot roll, pitch;
	float radius, uMul, vMul;
	float fog0, fog1;
};

__device__ __forceinline__ void tunnel_sample(const uint32_t *__restrict__ tex, int fpU, int fpV, float out[4])
{
	// bsamp_prepUVs(…, 1023, 10, …) + bsamp32_32f, bilinear.h:10-31, 86-135
	const unsigned U0 = unsigned(fpU >> 8), V0 = unsigned(fpV >> 8);
	const unsigned u0 = U0 & 1023, u1 = (U0+1) & 1023;
	const unsigned v0 = (V0 & 1023) << 10, v1 = ((V0+1) & 1023) << 10;
	const uint32_t s = bilerp_argb(__ldg(tex + u0 + v0), __ldg(tex + u1 + v0), __ldg(tex + u0 + v1), __ldg(tex + u1 + v1), fpU & 0xff, fpV & 0xff);
	out[0] = float(s & 0xff); out[1] = float((s >> 8) & 0xff); out[2] = float((s >> 16) & 0xff); out[3] = float(s >> 24);
}

__global__ void __launch_bounds__(kTileX*kTileY) tunnel_kernel(const TunnelFrame f, uint32_t *__restrict__ pDest, uint32_t *__restrict__ pGlowDest,
	const uint32_t *__restrict__ tex, const uint32_t *__restrict__ texGlow, const FrameGeom geom, const float2 *__restrict__ g_lut2, const RsqrtTab rsqrt, const TileQueue queue)
{
	__shared__ float2 s_lut2[2048];
	if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0)
		*queue.nextCounter = 0;
	const CosLut lut = stage_cos_lut(s_lut2, g_lut2);

	unsigned iX, iY;
	while (next_tile(queue, iX, iY))
	{
		if (iX >= unsigned(geom.fxX) || iY >= unsigned(geom.fxY))
			continue;

		float u, v;
		to_uv_fxmap(geom, iX, iY, 2.f, u, v);
		vec3 dir = { u, v, 1.f };
		rot_x(f.pitch, dir.y, dir.z);
		rot_z(f.roll, dir.x, dir.y);
		fast_norm3(rsqrt, dir);

		float A = dir.x*dir.x + dir.y*dir.y;
		A += f.flowerScale*lutcosf(lut, atan2f_ref(dir.y, dir.x)*f.flowerFreq + f.flowerPhase);

		const float absX = fabsf(dir.x), absY = fabsf(dir.y);
		const float box = absX > absY ? absX : absY;
		A = smoothstepf(A, box, f.boxy);
		A += kEpsilon;
		A = 1.f/A;
		const float T = f.radius*A;
		const float T2 = T*0.912f;
		const float ix = dir.x*T, iy = dir.y*T, iz = dir.z*T;
		const float ix2 = dir.x*T2, iy2 = dir.y*T2, iz2 = dir.z*T2;

		const float U = atan2f_ref(iy, ix)/kPI;
		const float V = iz + f.timeSpeed;
		const float U2 = atan2f_ref(iy2, ix2)/kPI;
		const float V2 = iz2 + f.timeSpeed;

		const int fpU = ftofp24(U*f.uMul), fpV = ftofp24(V*f.vMul);
		const int fpU2 = ftofp24(U2*f.uMul), fpV2 = ftofp24(V2*f.vMul);

		const float shade = clampf(0.f, 1.f, 1.f-expf_ref(-0.006f*T*T));

		float color[4], glow[4];
		tunnel_sample(tex, fpU, fpV, color);
		tunnel_sample(texGlow, fpU2, fpV2, glow);

		uint32_t px = 0, gpx = 0;
		#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			px |= to_chan_noconv(vlerp(color[i], f.fog0, shade)) << (8*i);
			gpx |= to_chan_noconv(vlerp(glow[i], f.fog1, shade)) << (8*i);
		}

		const size_t index = size_t(iY)*geom.fxX + iX;
		pDest[index] = px;
		pGlowDest[index] = gpx;
	}
}

// -------------------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------------------

FrameGeom MakeGeom(const ckd_ctx *ctx)
{
	FrameGeom g;
	g.fxX = ctx->fxX;
	g.fxY = ctx->fxY;
	g.invFxX = 1.f/float(ctx->fxX);
	g.invFxY = 1.f/float(ctx->fxY);
	g.aspect = float(ctx->resY)/float(ctx->resX);
	g.oneOverAspect = 1.f/g.aspect;
	return g;
}

Rot MakeRot(const ckd_ctx *ctx, float angle)
{
	return { ckdh::lutcosf(ctx->h_cosLUT, angle), ckdh::lutsinf(ctx->h_cosLUT, angle) };
}

// the queue of tile rows [row0, row1) (row1 < 0: all of them)
TileQueue MakeQueue(ckd_ctx *ctx, const FrameGeom &geom, int row0 = 0, int row1 = -1)
{
	TileQueue q;
	q.tilesX = ckd_div_up(geom.fxX, kWarpTileX);
	const unsigned tileRows = ckd_div_up(geom.fxY, kWarpTileY);
	q.firstTile = unsigned(q.tilesX)*unsigned(row0);
	q.numTiles = unsigned(q.tilesX)*(row1 < 0 ? tileRows : unsigned(row1));
	q.counter = ctx->d_tileCounters + (ctx->tileLaunches & 1);
	q.nextCounter = ctx->d_tileCounters + ((ctx->tileLaunches + 1) & 1);
	ctx->tileLaunches++;
	return q;
}

// Host-side proof that every LUT angle of a frame stays below the range of the conversion-free lookup (CosLutFast:
// scaled angle < 2^23, i.e. |angle| < 25735.9 rad).  It exists for the four effects whose distance function is BOUNDED --
// a sum of table values, each within [-1, 1] -- so the march total, hence every sample position, is bounded by the step
// count alone; what remains are the frame's own parameters (time offsets, origins), which 'add' folds in.  The spikey
// variants march |p| - radius, which grows geometrically: the close and the specular-only variant stay in range within their
// step budgets (SpikeyFixedProof); on the distant one rays that miss do reach such angles (and the reference's aliased
// lookups there are part of the picture), so it always runs the exact kernel.
// The induction behind the bound: while every angle so far was in range, every table value so far is within [-1, 1] (a
// lerp between two entries), so the next position obeys the bound, so the next angle is in range.  NaN/inf parameters fail
// the comparison and select the exact kernel.
struct LutRangeProof
{
	double worst = 0.0;            // largest |angle| (radians) any lookup of the frame can see
	void add(double bound) { if (!(bound <= worst)) worst = bound; } // NaN-proof: a NaN bound sticks
	bool holds() const { return worst < 20000.0; }                    // 22 % below 25735.9; the bounds themselves are generous and float rounding is 1e-7
};

template <class Effect> int LaunchRaymarch(ckd_ctx *ctx, const Effect &effect, uint32_t *d_fxmap, const char *name, const LutRangeProof *proof = nullptr,
	int tileRow0 = 0, int tileRow1 = -1)
{
	static const bool forceExact = nullptr != getenv("CKD_EXACT_LUT"); // tests: run every frame through the exact kernel
	const FrameGeom geom = MakeGeom(ctx);
	const RsqrtTab rsqrt = { ctx->d_rsqrtTab, ctx->rsqrtLog2Bin };
	const TileQueue queue = MakeQueue(ctx, geom, tileRow0, tileRow1);
	const unsigned tiles = queue.numTiles - queue.firstTile;
	// a persistent grid: exactly the CTAs the device holds at once.  A CTA beyond that would only start when an earlier one has
	// found the queue empty -- to stage its 16 KB table and leave
	const bool fast = nullptr != proof && proof->holds() && !forceExact;
	static int residentPerSM[2] = { 0, 0 };
	if (0 == residentPerSM[fast])
	{
		int n = 0;
		if (fast) CKD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, raymarch_kernel<Effect, true>, kTileX*kTileY, 0));
		else CKD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, raymarch_kernel<Effect, false>, kTileX*kTileY, 0));
		residentPerSM[fast] = std::max(1, n);
	}
	const int blocks = int(std::max(1u, std::min<unsigned>(ckd_div_up(tiles, kTileY), unsigned(ctx->numSMs)*unsigned(residentPerSM[fast]))));
	ckd_prof_begin(ctx, name, 4.0*kWarpTileX*kWarpTileY*tiles);
	if (fast)
		raymarch_kernel<Effect, true><<<blocks, dim3(kTileX, kTileY), 0, ctx->stream>>>(effect, d_fxmap, geom, ctx->d_cosLUT2, rsqrt, queue);
	else
		raymarch_kernel<Effect, false><<<blocks, dim3(kTileX, kTileY), 0, ctx->stream>>>(effect, d_fxmap, geom, ctx->d_cosLUT2, rsqrt, queue);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

// raymarch into the FX map, then Fx_Blit_2x2 into d_dest: the tail of every effect without a post chain.  With a read-back
// armed (ckd_arm_readback) both run per band of tile rows and every finished band of output rows leaves for the host on the
// copy stream while the next band renders.  Output rows 2y, 2y+1 need FX rows y and y+1, so a band's blit stops one FX row
// short of what has been rendered; the last band takes the rest.
template <class Effect> int RaymarchAndBlit(ckd_ctx *ctx, const Effect &effect, const char *name, const LutRangeProof *proof, uint32_t *d_dest)
{
	uint32_t *d_fxmap = ctx->d_fxMap[0];
	void *h_dest = ctx->rbHost;
	ctx->rbHost = nullptr;
	if (nullptr == h_dest)
	{
		CKD_TRY(LaunchRaymarch(ctx, effect, d_fxmap, name, proof));
		return ckd_fx_blit_2x2(ctx, d_dest, d_fxmap);
	}

	const int bands = ctx->rbBands, halfY = ctx->fxY - 4;
	const int tileRows = int(ckd_div_up(ctx->fxY, kWarpTileY));
	int blitted = 0;
	for (int k = 0; k < bands; ++k)
	{
		const int row0 = tileRows*k/bands, row1 = tileRows*(k + 1)/bands;
		if (row1 > row0)
			CKD_TRY(LaunchRaymarch(ctx, effect, d_fxmap, name, proof, row0, row1));
		const int y1 = (k == bands-1) ? halfY : std::max(blitted, std::min(row1*kWarpTileY - 1, halfY));
		if (y1 > blitted)
		{
			CKD_TRY(ckd_fx_blit_2x2_rows(ctx, d_dest, d_fxmap, blitted, y1));
			const size_t offset = size_t(blitted)*2*ctx->resX, count = size_t(y1 - blitted)*2*ctx->resX;
			CKD_CUDA(cudaEventRecord(ctx->evBand[k], ctx->stream));
			CKD_CUDA(cudaStreamWaitEvent(ctx->copyStream, ctx->evBand[k], 0));
			CKD_CUDA(cudaMemcpyAsync(static_cast<uint32_t *>(h_dest) + offset, d_dest + offset, count*sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->copyStream));
			blitted = y1;
		}
	}
	ctx->rbIssued = true;
	return CKD_OK;
}

// |u| and |v| of to_uv_fxmap for any pixel of the map
double MaxU(const FrameGeom &g, double scale) { return 0.5*scale*double(g.oneOverAspect); }
double MaxV(double scale) { return 0.5*scale; }

void CopyColor(float *dst, const ckdh::vec4 &c, int n)
{
	const float src[4] = { c.x, c.y, c.z, c.w };
	for (int i = 0; i < n; ++i) dst[i] = src[i];
}

} // namespace

// Plasma_Draw, shadertoy.cpp:276-280
extern "C" int ckd_plasma_draw(ckd_ctx *ctx, const ckd_plasma_params *p, float time, uint32_t *d_dest)
{
	CKD_REQUIRE(ctx && p && d_dest, "null argument");
	const float *lut = ctx->h_cosLUT;

	PlasmaEffect fx;
	const ckdh::vec4 colMulA = ckdh::Desaturate(ckdh::MichielPal(lut, p->hue), p->desaturation);
	const ckdh::vec4 colMulB = ckdh::Desaturate(colMulA, 0.8f);
	CopyColor(fx.f.colMulA, colMulA, 3);
	CopyColor(fx.f.colMulB, colMulB, 3);
	time = time*p->speed;
	const float angle = time*0.314f*0.5f;
	fx.f.time = time;
	fx.f.dirCos = ckdh::lutcosf(lut, angle);
	fx.f.dirSin = ckdh::lutsinf(lut, angle);
	fx.f.gamma = p->gamma;

	// fPlasma = |(fX, fY, fZ)| - 0.8 with |f| <= 1.2: march in [-0.8, 1.28], 24 steps of march*0.809 -> |total| <= 24.9;
	// h = d*total with |dx| <= |u|*aspect + 0.75, |dy| <= |v|, |dz| <= |u| + 0.75; angles: hx - hy + pi/2, hx*0.33, hy*0.43,
	// (5*time + hz)*0.53 (shadertoy.cpp:201-209)
	LutRangeProof proof;
	{
		const FrameGeom g = MakeGeom(ctx);
		const double total = 25.0, u = MaxU(g, 4.0), v = MaxV(4.0);
		const double hx = (u*double(g.aspect) + 0.75)*total, hy = v*total, hz = (u + 0.75)*total;
		proof.add(hx + hy + 1.6);
		proof.add((5.0*fabs(double(time)) + hz)*0.53);
	}
	return RaymarchAndBlit(ctx, fx, "raymarch_plasma", &proof, d_dest);
}

// Nautilus_Draw, shadertoy.cpp:395-407
extern "C" int ckd_nautilus_draw(ckd_ctx *ctx, const ckd_nautilus_params *p, float time, uint32_t *d_dest)
{
	CKD_REQUIRE(ctx && p && d_dest, "null argument");
	const float *lut = ctx->h_cosLUT;

	NautilusEffect fx;
	time = time*p->speed;
	fx.f.time = time;
	fx.f.gx = time*0.125f;
	fx.f.gy = time/9.f;
	fx.f.gz = ckdh::lutcosf(lut, time*0.1428f);
	const ckdh::vec4 colorization = { .1f-ckdh::lutcosf(lut, p->hue/3.f)/19.f, .1f, .1f+ckdh::lutcosf(lut, p->hue/14.f)/8.f, 0.f };
	CopyColor(fx.f.diffColor, ckdh::Desaturate(colorization, p->desaturation), 3);
	fx.f.cosHitOffs = ckdh::lutcosf(lut, time*0.314f*0.5f);
	fx.f.funkCos = ckdh::lutcosf(lut, time*ckdh::kGoldenRatio*0.1f);
	fx.f.roll = MakeRot(ctx, p->roll*time);

	// fNautilus = dotted*0.5 - 0.7 in [-0.7, 0.8]; total starts at 0.01 and takes <= 48 steps of march*0.628 -> |total| <= 24.2;
	// |dir| = 1 (+ the RSQRTPS error), taps at +0.15 and +cosHitOffs: |p| <= 24.3 + 0.15 + |cosHitOffs| (shadertoy.cpp:289-298)
	LutRangeProof proof;
	{
		const double p = 24.3*1.001 + 0.15 + fabs(double(fx.f.cosHitOffs));
		proof.add(p + fabs(double(fx.f.gx)));
		proof.add(p + fabs(double(fx.f.gy)));
		proof.add(2.0*1.001*p);                                  // cos*px - cos*py
		proof.add(0.33*p*p + fabs(double(fx.f.gz))*p);
		proof.add(2.8*p + fabs(double(fx.f.time)));
	}
	const float blur = ckdh::BoxBlurScale(p->blur);
	if (0.f == blur)
		return RaymarchAndBlit(ctx, fx, "raymarch_nautilus", &proof, d_dest);
	ctx->rbHost = nullptr; // the blur works on the whole frame: no banded read-back
	CKD_TRY(LaunchRaymarch(ctx, fx, ctx->d_fxMap[0], "raymarch_nautilus", &proof));
	CKD_TRY(ckd_fx_blit_2x2(ctx, d_dest, ctx->d_fxMap[0]));
	return ckd_old_blur(ctx, d_dest, d_dest, unsigned(ctx->resX), unsigned(ctx->resY), blur);
}

// Spikey_Draw, shadertoy.cpp:661-733
// Range proof of the close and the specular-only variants (fixed step budgets, slow growth).  March: |fSpikey| <= |p| + 1.35 +
// 2*scale while every table value so far is within [-1, 1] (the induction of LutRangeProof); |p| <= |origin| + |dir|*|total|
// with |dir| <= 1.002 (RSQRTPS error); total grows by |march|*stepScale per step.  The recursion is evaluated in double with
// every constant rounded up; the taps add nOffs.  Close: 32 steps of 0.157 -> |total| <= ~270 for the shipped parameters,
// angles <= 22*272 + |gx|: far inside the range.  Anything non-finite fails LutRangeProof::holds().
static LutRangeProof SpikeyFixedProof(const SpikeyFrame &f, double originLen, double stepScale, int steps, double nOffs)
{
	double total = 0.0;
	for (int i = 0; i < steps; ++i)
		total += stepScale*1.001*((originLen + 1.002*total) + 1.35 + 0.5);
	const double pMax = originLen + 1.002*total + fabs(nOffs);
	LutRangeProof proof;
	proof.add(fabs(double(f.gy))*pMax + fabs(double(f.gx)));
	proof.add(fabs(double(f.gz))*pMax + fabs(double(f.gx)));
	return proof;
}

extern "C" int ckd_spikey_draw(ckd_ctx *ctx, const ckd_spikey_params *p, float time, int close, uint32_t *d_dest)
{
	CKD_REQUIRE(ctx && p && d_dest, "null argument");
	const float *lut = ctx->h_cosLUT;
	const unsigned fxSize = unsigned(ctx->fxX)*unsigned(ctx->fxY);
	const float aspect = float(ctx->resY)/float(ctx->resX);

	SpikeyFrame f = {};
	f.roll = MakeRot(ctx, p->roll);
	f.specPow = p->specular;
	f.gamma = p->gamma;
	CopyColor(f.diffColor, ckdh::Desaturate(ckdh::MichielPal(lut, p->hue), p->desaturation), 4);

	if (close)
	{
		CKD_REQUIRE(ctx->images[CKD_IMG_SPIKE_BLUR_MAP0].d_pixels && ctx->images[CKD_IMG_SPIKE_BLUR_MAP1].d_pixels, "spike blur maps not uploaded");

		// RenderSpikeyMap_2x2_Close, shadertoy.cpp:432-458
		const float zOffsFinal = ckdh::easeInOutElasticf(p->close_z)*p->close_z_scale;
		f.gx = p->speed*time;
		f.gy = 16.f*p->close_scale;
		f.gz = (0 == p->close_aspect_mul) ? 22.f*p->close_scale : aspect*22.f*p->close_scale;
		f.xOffs = p->close_x;
		f.yOffs = p->close_y;
		f.zTerm = 1.f + zOffsFinal;
		f.normalGrain = p->close_normal_grain;
		SpikeyCloseEffect fx = { f };
		// origin (0.2, 0, -2.23), 32 steps of march*(0.05*pi), taps at +normalGrain (shadertoy.cpp:460-492)
		const LutRangeProof proof = SpikeyFixedProof(f, 2.24, 0.05*3.14159266, 32, double(f.normalGrain));
		const float mbOpacity = ckdh::saturatef(p->mix_blur_opacity);
		if (!(mbOpacity > 0.f))
			return RaymarchAndBlit(ctx, fx, "raymarch_spikey_close", &proof, d_dest);
		CKD_TRY(LaunchRaymarch(ctx, fx, ctx->d_fxMap[0], "raymarch_spikey_close", &proof));
		{
			// shadertoy.cpp:672-707
			const float mbMap = ckdh::clampf(0, 1.f, p->mix_blur_map);
			const float mbBlur = ckdh::clampf(0.f, 100.f, p->mix_blur);
			const float mbMapBlur = ckdh::clampf(0.f, 100.f, p->mix_map_blur);
			const size_t fxBytes = size_t(fxSize)*4;
			const void *map0 = ctx->images[CKD_IMG_SPIKE_BLUR_MAP0].d_pixels, *map1 = ctx->images[CKD_IMG_SPIKE_BLUR_MAP1].d_pixels;

			if (0.f == mbMap)
				CKD_CUDA(cudaMemcpyAsync(ctx->d_spikeBlurMap, map0, fxBytes, cudaMemcpyDeviceToDevice, ctx->stream));
			else if (1.f == mbMap)
				CKD_CUDA(cudaMemcpyAsync(ctx->d_spikeBlurMap, map1, fxBytes, cudaMemcpyDeviceToDevice, ctx->stream));
			else
			{
				CKD_CUDA(cudaMemcpyAsync(ctx->d_spikeBlurMap, map0, fxBytes, cudaMemcpyDeviceToDevice, ctx->stream));
				CKD_TRY(ckd_blend(ctx, CKD_MIX32, ctx->d_spikeBlurMap, static_cast<const uint32_t *>(map1), fxSize, 0.f, ckdh::x86_f2u(mbMap*255.f) & 0xff));
			}

			CKD_CUDA(cudaMemcpyAsync(ctx->d_fxMap[1], ctx->d_fxMap[0], fxBytes, cudaMemcpyDeviceToDevice, ctx->stream));

			if (mbMapBlur >= 1.f)
				CKD_TRY(ckd_old_blur(ctx, ctx->d_spikeBlurMap, ctx->d_spikeBlurMap, unsigned(ctx->fxX), unsigned(ctx->fxY), ckdh::BoxBlurScale(mbMapBlur)));
			if (mbBlur >= 1.f)
				CKD_TRY(ckd_old_blur(ctx, ctx->d_fxMap[1], ctx->d_fxMap[1], unsigned(ctx->fxX), unsigned(ctx->fxY), ckdh::BoxBlurScale(mbBlur)));

			CKD_TRY(ckd_blend(ctx, CKD_SOFTLIGHT32AA, ctx->d_fxMap[1], ctx->d_spikeBlurMap, fxSize, tanhf(mbBlur+mbOpacity), 0));
			CKD_TRY(ckd_blend(ctx, CKD_OVERLAY32A, ctx->d_fxMap[0], ctx->d_fxMap[1], fxSize, 0.f, 0));
		}
		return ckd_fx_blit_2x2(ctx, d_dest, ctx->d_fxMap[0]);
	}

	const float warmup = p->warmup;
	if (0.f == warmup)
	{
		// RenderSpikeyMap_2x2_Distant, shadertoy.cpp:525-544
		f.gx = p->speed*time; f.gy = 16.f; f.gz = 16.f;
		f.xOffs = p->dist_x;
		f.yOffs = p->dist_y;
		f.zTerm = -2.614f + p->dist_z;
		// 48 steps of 0.314*(|p| - radius): rays that miss the ball run away geometrically (x 1.314 per step) and do reach angles
		// past the fast lookup's range -- the reference's aliased lookups out there are part of the picture -- so this variant
		// always runs the exact kernel.  (A per-evaluation guard on the march total that switches lookups inside the FAST
		// kernel was measured: 186.6 us against 175.7 us at 4K, profiles/r02_notes.md.)
		SpikeyDistantEffect fx = { f };
		return RaymarchAndBlit(ctx, fx, "raymarch_spikey_distant", nullptr, d_dest);
	}

	// RenderSpikeyMap_2x2_Distant_SpecularOnly(…, 1.f+warmup), shadertoy.cpp:600-608, 727-729
	f.gx = p->speed*time; f.gy = 8.f; f.gz = 16.f;
	f.warmup = 1.f+warmup;
	SpikeySpecOnlyEffect fx = { f };
	// origin (0, 0, -3.314), 36 steps of march*0.075*golden ratio, taps at +0.01 (shadertoy.cpp:610-640)
	const LutRangeProof proof = SpikeyFixedProof(f, 3.314, 0.075*1.6180340, 36, 0.01);
	CKD_TRY(LaunchRaymarch(ctx, fx, ctx->d_fxMap[0], "raymarch_spikey_spec", &proof));
	CKD_TRY(ckd_old_blur_h(ctx, ctx->d_fxMap[0], ctx->d_fxMap[0], unsigned(ctx->fxX), unsigned(ctx->fxY), ckdh::BoxBlurScale(1.f+warmup)));
	return ckd_fx_blit_2x2(ctx, d_dest, ctx->d_fxMap[0]);
}

// Tunnel_Draw, shadertoy.cpp:840-861
extern "C" int ckd_tunnel_draw(ckd_ctx *ctx, const ckd_tunnel_params *p, float time, uint32_t *d_dest)
{
	CKD_REQUIRE(ctx && p && d_dest, "null argument");
	const ckd_image_slot &tex = ctx->images[CKD_IMG_TUNNEL_TEX], &texFx = ctx->images[CKD_IMG_TUNNEL_TEX_FX];
	if (!tex.d_pixels || !texFx.d_pixels || tex.width != 1024 || tex.height != 1024 || texFx.width != 1024 || texFx.height != 1024 || tex.bpp != 4 || texFx.bpp != 4)
	{
		ckd_set_error("ckd_tunnel_draw: the two 1024x1024 BGRA tunnel textures have not been uploaded (shadertoy.cpp:174-175)");
		return CKD_ERR_MISSING_INPUT;
	}

	TunnelFrame f;
	f.boxy = p->boxy;
	f.flowerScale = p->flower_scale;
	f.flowerFreq = p->flower_freq;
	f.flowerPhase = p->flower_phase*time;
	f.roll = MakeRot(ctx, p->roll*time);
	f.pitch = MakeRot(ctx, p->pitch*time);
	f.radius = p->radius;
	f.uMul = p->mul_u;
	f.vMul = p->mul_v;
	f.fog0 = p->fog1;
	f.fog1 = p->fog2;
	time *= p->speed;
	f.timeSpeed = time*p->speed;

	const FrameGeom geom = MakeGeom(ctx);
	const TileQueue queue = MakeQueue(ctx, geom);
	const int blocks = int(std::min<unsigned>(ckd_div_up(queue.numTiles, kTileY), unsigned(ctx->numSMs)*8));
	const RsqrtTab rsqrt = { ctx->d_rsqrtTab, ctx->rsqrtLog2Bin };
	ckd_prof_begin(ctx, "raymarch_tunnel", 8.0*geom.fxX*geom.fxY);
	tunnel_kernel<<<blocks, dim3(kTileX, kTileY), 0, ctx->stream>>>(f, ctx->d_fxMap[0], ctx->d_fxMap[1],
		static_cast<const uint32_t *>(tex.d_pixels), static_cast<const uint32_t *>(texFx.d_pixels), geom, ctx->d_cosLUT2, rsqrt, queue);
	CKD_CHECK_LAUNCH(ctx);

	const float litBlur = ckdh::clampf(0.f, 100.f, p->lit_blur);
	if (0 != p->lit_tiles)
	{
		const unsigned fxSize = unsigned(ctx->fxX)*unsigned(ctx->fxY);
		if (litBlur >= 1.f)
			CKD_TRY(ckd_old_blur(ctx, ctx->d_fxMap[1], ctx->d_fxMap[1], unsigned(ctx->fxX), unsigned(ctx->fxY), ckdh::BoxBlurScale(litBlur)));
		CKD_TRY(ckd_blend(ctx, CKD_ADD32, ctx->d_fxMap[0], ctx->d_fxMap[1], fxSize, 0.f, 0));
	}
	return ckd_fx_blit_2x2(ctx, d_dest, ctx->d_fxMap[0]);
}

// Sinuses_Draw, shadertoy.cpp:984-988
extern "C" int ckd_sinuses_draw(ckd_ctx *ctx, const ckd_sinuses_params *p, float time, uint32_t *d_dest)
{
	CKD_REQUIRE(ctx && p && d_dest, "null argument");
	const float *lut = ctx->h_cosLUT;

	SinusesEffect fx;
	fx.f.specPow = 1.f + p->specular;
	fx.f.roll = MakeRot(ctx, p->roll);
	fx.f.offsX = p->offs_x;
	fx.f.gamma = p->gamma;
	CopyColor(fx.f.diffColor, ckdh::Desaturate(ckdh::MichielPal(lut, p->hue), p->desaturation), 4);

	// fSinPath(time*speed), shadertoy.cpp:870-876
	const float pathTime = time*p->speed;
	const float timeMod = pathTime*0.314f;
	const float sine = ckdh::lutsinf(lut, timeMod);
	const float cosine = ckdh::lutcosf(lut, timeMod);
	fx.f.origin[0] = sine*2.f*ckdh::kGoldenRatio - cosine*1.5f;
	fx.f.origin[1] = cosine*3.14f + sine*ckdh::kGoldenRatio;
	fx.f.origin[2] = pathTime;

	// fSinMap = (|(cosX, cosY, cosZ)| - 1.025)*1.33 in [-1.37, 0.95]; <= 32 steps of march*0.814 -> |total| <= 35.7; |dir| = 1;
	// normal taps at +0.2: |p_i| <= |origin_i| + 36; inside: |pX| <= |px0| + 4.74, |pY| <= |py0| + 4.76, every angle is at
	// most 1.0175*|coordinate| + pi/2 or 0.394*|coordinate| + 1 (shadertoy.cpp:879-899)
	LutRangeProof proof;
	for (int i = 0; i < 3; ++i)
		proof.add(1.0175*(fabs(double(fx.f.origin[i])) + 36.0*1.001 + 0.2 + 4.76) + 1.6);
	return RaymarchAndBlit(ctx, fx, "raymarch_sinuses", &proof, d_dest);
}

// Laura_Draw, shadertoy.cpp:1105-1109
extern "C" int ckd_laura_draw(ckd_ctx *ctx, const ckd_laura_params *p, float time, uint32_t *d_dest)
{
	CKD_REQUIRE(ctx && p && d_dest, "null argument");
	const float *lut = ctx->h_cosLUT;

	LauraEffect fx;
	CopyColor(fx.f.diffColor, ckdh::Desaturate(ckdh::MichielPal(lut, p->hue), p->saturate), 4);
	fx.f.originZ = p->speed*time;
	fx.f.yaw = MakeRot(ctx, p->yaw);
	fx.f.pitch = MakeRot(ctx, p->pitch);
	fx.f.roll = MakeRot(ctx, p->roll*time);

	// fLaura = three table values + 1 in [-2, 4]; 32 steps of march*0.5 -> |total| <= 64; |dir| = 1; hit = origin + dir*total,
	// normal taps at +0.1628; the angles are the coordinates themselves (shadertoy.cpp:998-1015)
	LutRangeProof proof;
	proof.add(fabs(double(fx.f.originZ)) + 64.1*1.001 + 0.17);
	return RaymarchAndBlit(ctx, fx, "raymarch_laura", &proof, d_dest);
}
