/*
 * ckd.h -- C ABI of cookiedough_b200: the B200 (sm_100a) implementation of cookiedough's per-pixel effect hot path.
 *
 * Every entry point is plain C (opaque context, POD parameter structs, raw device/host pointers and sizes) so the
 * reference -- or any FFI -- can bind it directly.  Each declaration cites the reference interface it replaces
 * (paths relative to the reference's code/ directory).  The C++ host layer (include/ckd_host.h) re-creates the
 * reference's own X_Create / X_Draw(uint32_t *pDest, float time, float delta) / X_Destroy names on top of this ABI.
 *
 * Conventions
 *   - pixels are uint32_t ARGB8888 (0xAARRGGBB little endian = bytes B,G,R,A), row-major, no padding (code/main.h:37-43)
 *   - "d_" pointers are device addresses on the context's GPU; they may alias where the reference allows in-place use
 *   - all work is enqueued on the context's stream (ckd_set_stream); nothing synchronises unless documented
 *   - every function returns CKD_OK (0) or a negative ckd_status; ckd_last_error() describes the last failure
 *   - there is NO CPU fallback: without a usable CUDA device ckd_create() fails with CKD_ERR_CUDA
 */
#ifndef CKD_H_
#define CKD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ckd_ctx ckd_ctx;

typedef enum ckd_status {
	CKD_OK = 0,
	CKD_ERR_CUDA = -1,          /* CUDA runtime/driver error (message in ckd_last_error) */
	CKD_ERR_INVALID = -2,       /* bad argument */
	CKD_ERR_MISSING_INPUT = -3, /* an image / table the effect needs has not been uploaded */
	CKD_ERR_UNIMPLEMENTED = -4,
	CKD_ERR_TIMEOUT = -5        /* a device-side wait of the frame gather gave up (ckd_gather_status) */
} ckd_status;

/* ---- context (replaces the module globals allocated by Shared_Create / FxBlitter_Create / Polar_Create /
 *      BoxBlur_Create: shared-resources.cpp:14-37, fx-blitter.cpp:10-20, polar.cpp:61-72, boxblur.cpp:23-28;
 *      and CalculateCosLUT / InitializeFastCosine: sincos-lut.cpp:9-16, fast-cosine.cpp:10-17) ------------------- */

/* resX/resY replace the compile-time kResX/kResY (main.h:37-38); both must be multiples of 8.
 * ckd_create makes `device` the calling thread's current CUDA device and the context lives there.  The other ckd_* calls do
 * not switch devices (they are launch-rate sensitive): a process drives ONE device -- the deployment model is one process
 * per GPU (bench.py, tools/render_demo.py) -- or the caller makes the context's device current (cudaSetDevice) before using
 * a context that lives on another one.  ckd_last_error() returns the calling thread's latest failure (or, when this thread
 * has had none, the latest failure of any thread). */
int ckd_create(ckd_ctx **out_ctx, int res_x, int res_y, int device);
void ckd_destroy(ckd_ctx *ctx);
const char *ckd_last_error(void);
const char *ckd_version(void);

int ckd_set_stream(ckd_ctx *ctx, void *cuda_stream); /* cudaStream_t; NULL = default stream */
/* second context on the same device for frames in flight side by side (two frames of a timeline overlap their latency-bound
 * kernels: blurs, casters): ckd_own_stream gives the context a non-blocking stream of its own (destroyed with the context),
 * ckd_clone_inputs copies everything a context is configured with -- the uploaded images, the cosine / fast-cosine / RSQRTPS
 * tables, the polar maps, the frame-independent flag -- from `src` into `dst` (same resolution, same device), and ckd_join makes
 * `ctx`'s stream wait for everything enqueued on `other`'s stream so far. */
int ckd_own_stream(ckd_ctx *ctx);
int ckd_clone_inputs(ckd_ctx *dst, const ckd_ctx *src);
int ckd_join(ckd_ctx *ctx, ckd_ctx *other);
int ckd_sync(ckd_ctx *ctx);

int ckd_res_x(const ckd_ctx *ctx);
int ckd_res_y(const ckd_ctx *ctx);
int ckd_fxmap_res_x(const ckd_ctx *ctx); /* kFxMapResX = resX/2+4 (fx-blitter.h:16) */
int ckd_fxmap_res_y(const ckd_ctx *ctx);

/* device twins of the reference's global buffers */
uint32_t *ckd_frame(ckd_ctx *ctx);                  /* device copy of pDest (main.cpp:307) */
uint32_t *ckd_fxmap(ckd_ctx *ctx, int index);       /* g_pFxMap[0..3] (fx-blitter.h:26) */
uint32_t *ckd_render_target(ckd_ctx *ctx, int index); /* g_renderTarget[0..3] (shared-resources.h:12) */

/* generic device memory + copies (host pointers may be pageable or pinned) */
int ckd_malloc(ckd_ctx *ctx, void **out_d_ptr, size_t bytes);
int ckd_free(ckd_ctx *ctx, void *d_ptr);
int ckd_malloc_host(void **out_h_ptr, size_t bytes); /* pinned */
int ckd_free_host(void *h_ptr);
int ckd_pin_host(void *h_ptr, size_t bytes);         /* page-lock a caller-owned buffer in place (e.g. the reference's pDest, main.cpp:307) */
int ckd_unpin_host(void *h_ptr);
int ckd_upload(ckd_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);   /* async on the stream */
int ckd_download(ckd_ctx *ctx, void *h_dst, const void *d_src, size_t bytes); /* async on the stream */
int ckd_copy(ckd_ctx *ctx, void *d_dst, const void *d_src, size_t bytes);       /* device to device, async on the stream (the compositor's memcpy, demo.cpp:407-449) */

/* banded read-back for synchronous callers: arms the NEXT effect draw with the caller's page-locked frame buffer.  A draw that
 * ends in raymarch + Fx_Blit_2x2 (Plasma, Sinuses, Laura, Spikey without its mix-blur chain, Nautilus without blur) or in a
 * polar remap (Tunnelscape without blur, Ball, Twister) then issues those stages in `bands` row bands and copies each finished
 * band to h_dest while the next one renders; every other draw, and
 * pageable h_dest, leave the arm unused.  ckd_finish_readback waits for the band copies and sets *out_done to 1 when the
 * frame is in h_dest, to 0 when the caller still has to copy it (ckd_download). */
int ckd_arm_readback(ckd_ctx *ctx, void *h_dest, int bands);
int ckd_finish_readback(ckd_ctx *ctx, int *out_done);

/* overlapped read-back for frame pipelines: the copy of a finished frame runs on the context's copy stream while the
 * compute stream already renders the next one.  slot = 0/1 selects one of two in-flight copies; ckd_frame_slot(ctx, slot)
 * are two device frame buffers to alternate between (images of their own: no effect, blur or gather uses them as scratch).  ckd_download_overlapped() orders the copy after everything enqueued
 * on the compute stream so far; ckd_wait_download() blocks the host until that slot's copy has landed. */
uint32_t *ckd_frame_slot(ckd_ctx *ctx, int slot);
int ckd_download_overlapped(ckd_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, int slot);
int ckd_wait_download(ckd_ctx *ctx, int slot);

/* timing on the context's stream (CUDA events) */
int ckd_timer_start(ckd_ctx *ctx);
int ckd_timer_stop_ms(ckd_ctx *ctx, float *out_ms); /* synchronises */

/* ---- tables ---------------------------------------------------------------------------------------------------
 * ckd_create() fills all of these itself with the host's libm / CPU exactly like the reference does at start-up;
 * the setters exist so tests can pin tables captured on another machine. */
int ckd_set_cos_lut(ckd_ctx *ctx, const float *lut2049);                      /* g_cosLUT (sincos-lut.cpp:6) */
/* _mm_rsqrt_ps emulation (shadertoy-util.h:89,260,265): table[parity][mantissa >> log2_bin] holds the result bits for
 * x = 2^(parity-1) * 1.mantissa, parity 0: [0.5,1), parity 1: [1,2); entries = 2 << (23 - log2_bin). */
int ckd_set_rsqrt_table(ckd_ctx *ctx, const uint32_t *table, int log2_bin);
int ckd_get_rsqrt_table(ckd_ctx *ctx, uint32_t *out_table, size_t max_entries, int *out_log2_bin, size_t *out_entries);
int ckd_set_polar_maps(ckd_ctx *ctx, const int32_t *map, const int32_t *inv_map); /* s_pMap/s_pInvMap (polar.cpp:13-14), 2 ints/px */
int ckd_get_polar_maps(ckd_ctx *ctx, int32_t *out_map, int32_t *out_inv_map);

/* g_fastCosTab (fast-cosine.cpp:11-17): 1025 doubles cos(i*2pi/1024), built by ckd_create with the host's cos() like
 * InitializeFastCosine does; the getter returns the host copy. */
int ckd_set_fast_cos_table(ckd_ctx *ctx, const double *table1025);
int ckd_get_fast_cos_table(ckd_ctx *ctx, double *out_table1025);
/* fastcosf (fast-cosine.h:17-49) / fastsinf (:51-53) over n arguments on the device: d_out[i] = fastcosf(d_x[i]).  The table is
 * staged in shared memory once per CTA.  No effect of the demo calls fastcosf; the entry exists so the helper is available
 * (and testable) on the device side of the boundary like lutcosf is. */
int ckd_fastcos(ckd_ctx *ctx, float *d_out, const double *d_x, size_t n, int sine);

/* Frame independence.  Ball_Draw with beams leaves the last pixel of every ray row of g_renderTarget[0] untouched
 * (ball.cpp:168-203 fills up to kTargetResX-1, and :352-363 does not clear), so that column -- and, through the in-place
 * blur, its neighbourhood -- carries over from whatever frame was rendered before.  With enabled != 0 that column is cleared
 * to 0 before the rays are cast, which makes every frame a pure function of (time, tracks, assets): required when frames
 * are sharded over GPUs, where "the frame before" differs with the GPU count.  Default 0 = the reference's behaviour. */
int ckd_set_frame_independent(ckd_ctx *ctx, int enabled);

/* ---- images (replace Image_Load32 / Image_Load8 results held in file statics) -------------------------------------- */
typedef enum ckd_image {
	CKD_IMG_TUNNEL_TEX = 0,      /* shadertoy.cpp:174  1024x1024 BGRA */
	CKD_IMG_TUNNEL_TEX_FX,       /* shadertoy.cpp:175  1024x1024 BGRA */
	CKD_IMG_SPIKE_BLUR_MAP0,     /* shadertoy.cpp:180  FX-map sized BGRA */
	CKD_IMG_SPIKE_BLUR_MAP1,     /* shadertoy.cpp:181 */
	CKD_IMG_SCAPE_HEIGHT,        /* landscape.cpp:200  1024x1024 L8 */
	CKD_IMG_SCAPE_COLOR,         /* landscape.cpp:201  1024x1024 BGRA */
	CKD_IMG_SCAPE_FOG,           /* landscape.cpp:206  256x1 BGRA */
	CKD_IMG_TSCAPE_HEIGHT,       /* tunnelscape.cpp:139 2048x2048 L8 */
	CKD_IMG_TSCAPE_COLOR,        /* tunnelscape.cpp:140 2048x2048 BGRA */
	CKD_IMG_TSCAPE_FOG,          /* tunnelscape.cpp:145 256x1 BGRA */
	CKD_IMG_BALL_HEIGHT0,        /* ball.cpp:367-380 (5 maps, 1024x1024 L8) */
	CKD_IMG_BALL_HEIGHT1,
	CKD_IMG_BALL_HEIGHT2,
	CKD_IMG_BALL_HEIGHT3,
	CKD_IMG_BALL_HEIGHT4,
	CKD_IMG_BALL_COLOR0,         /* ball.cpp:394-395 */
	CKD_IMG_BALL_COLOR1,
	CKD_IMG_BALL_BEAM0,          /* ball.cpp:400-402 */
	CKD_IMG_BALL_BEAM1,
	CKD_IMG_BALL_BEAM2,
	CKD_IMG_BALL_ENV,            /* ball.cpp:407 */
	CKD_IMG_BALL_BACKGROUND0,    /* ball.cpp:412-413 output sized */
	CKD_IMG_BALL_BACKGROUND1,
	CKD_IMG_BALL_HALO,           /* ball.cpp:422 output sized */
	CKD_IMG_TWISTER_HEIGHT,      /* torus-twister.cpp:143 1024x1024 L8 */
	CKD_IMG_TWISTER_COLOR,       /* torus-twister.cpp:144 */
	CKD_IMG_TWISTER_BACKGROUND,  /* torus-twister.cpp:149 output sized */
	CKD_IMG_COUNT
} ckd_image;

/* bytes_per_pixel: 1 (L8) or 4 (BGRA).  The pixels are copied to the device. */
int ckd_set_image(ckd_ctx *ctx, ckd_image slot, const void *h_pixels, int width, int height, int bytes_per_pixel);
const void *ckd_get_image(ckd_ctx *ctx, ckd_image slot); /* device pointer or NULL */

/* ---- 2D post chain ---------------------------------------------------------------------------------------------- */

/* Fx_Blit_2x2(pDest, pSrc) fx-blitter.cpp:27-75: FX map (fxResX x fxResY) -> output (resX x resY) */
int ckd_fx_blit_2x2(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src);

/* Polar_Blit / Polar_BlitA (polar.cpp:135-154, 180-198) */
int ckd_polar_blit(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int inverse);
int ckd_polar_blit_a(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int inverse);
/* Polar_Blit_2x2 (polar.cpp:200-218): FX-map sized source and destination (fxResX x fxResY), FX-map sized maps built on
 * first use.  The reference's 64/32-pixel tiles do not divide the FX map and run past the end of its buffers; this
 * entry writes exactly fxResX*fxResY pixels (the in-bounds part of the reference's result). */
int ckd_polar_blit_2x2(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int inverse);

/* 2007 box blur (deprecated/boxblur.cpp:44-231); d_dest may equal d_src (the reference's in-place semantics are kept) */
int ckd_old_blur_h(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength);
int ckd_old_blur_v(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength);
int ckd_old_blur(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength);
float ckd_box_blur_scale(float strength); /* BoxBlurScale, deprecated/boxblur.h:13-19 */

/* 2026 multi-pass blur (boxblur.cpp:270-318).  Rows are read 2 pixels past their end like the reference does
 * (boxblur.cpp:171,185): d_src must have >= 2 readable pixels after the last row. */
int ckd_new_blur_h(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float gain, unsigned num_passes);
int ckd_new_blur_v(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float gain, unsigned num_passes);
int ckd_new_blur(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float gain, unsigned num_passes);

/* blend ops on equally sized buffers (util.cpp:83-812, util.h:73-122) */
typedef enum ckd_blend_op {
	CKD_MIX32 = 0,        /* Mix32(dest, src, n, alpha)            u_param = alpha (uint8)      util.cpp:83  */
	CKD_MIXOVER32 = 1,    /* MixOver32                                                          util.cpp:148 */
	CKD_ADD32 = 2,        /* Add32                                                              util.cpp:132 */
	CKD_SUB32 = 3,        /* Sub32                                                              util.cpp:179 */
	CKD_EXCL32 = 4,       /* Excl32                                                             util.cpp:194 */
	CKD_SOFTLIGHT32 = 5,  /* SoftLight32                                                        util.cpp:242 */
	CKD_SOFTLIGHT32A = 6, /* SoftLight32A                                                       util.cpp:274 */
	CKD_SOFTLIGHT32AA = 7,/* SoftLight32AA(dest, src, n, alpha)    f_param = alpha              util.cpp:309 */
	CKD_OVERLAY32 = 8,    /* Overlay32                                                          util.cpp:434 */
	CKD_OVERLAY32A = 9,   /* Overlay32A                                                         util.cpp:485 */
	CKD_DARKEN32_50 = 10, /* Darken32_50                                                        util.cpp:518 */
	CKD_MULSRC32 = 11,    /* MulSrc32                                                           util.cpp:605 */
	CKD_MULSRC32A = 12,   /* MulSrc32A                                                          util.cpp:620 */
	CKD_MIXSRC32 = 13,    /* MixSrc32                                                           util.cpp:661 */
	CKD_FADE32 = 14       /* Fade32(dest, n, RGB, alpha)  u_param = alpha<<24 | RGB; src unused util.cpp:798 */
} ckd_blend_op;
int ckd_blend(ckd_ctx *ctx, ckd_blend_op op, uint32_t *d_dest, const uint32_t *d_src, unsigned num_pixels, float f_param, unsigned u_param);
/* the same result as num_steps ckd_blend calls on d_dest, applied per pixel in one pass over the frame (the compositor's
 * layer stacks, demo.cpp:511-1000).  A step whose source is d_dest itself reads the running pixel. */
typedef struct ckd_blend_step { ckd_blend_op op; const uint32_t *d_src; float f_param; unsigned u_param; } ckd_blend_step;
int ckd_blend_chain(ckd_ctx *ctx, uint32_t *d_dest, const ckd_blend_step *steps, unsigned num_steps, unsigned num_pixels);

/* rectangular blits (util.cpp:707-796) */
typedef enum ckd_blit_op {
	CKD_BLITSRC32 = 0, CKD_BLITSRC32A = 1, CKD_BLITADD32 = 2, CKD_BLITADD32A = 3
} ckd_blit_op;
int ckd_blit(ckd_ctx *ctx, ckd_blit_op op, uint32_t *d_dest, const uint32_t *d_src, unsigned dest_res_x, unsigned src_res_x, unsigned y_res, float alpha);
/* MixSrc32S (util.cpp:636-659) */
int ckd_mix_src_s(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned dest_res_x, unsigned dest_res_y, unsigned src_stride);

/* memset32 (util.h:57-67) and TapeWarp32 (util.cpp:552-603; note the reference's sampler row stride is resX, not x_res) */
int ckd_memset32(ckd_ctx *ctx, uint32_t *d_dest, uint32_t value, size_t num_ints);
int ckd_tape_warp(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float speed);

/* ---- effects: one entry per reference X_Draw; the structs carry every Rocket-derived scalar the reference pulls
 *      with Rocket::getf/geti inside the call (track names in comments) ------------------------------------------- */

typedef struct ckd_plasma_params {       /* shadertoy.cpp:213-216 */
	float speed;         /* plasma:Speed */
	float hue;           /* plasma:Hue */
	float gamma;         /* plasma:Gamma */
	float desaturation;  /* plasma:Desaturation */
} ckd_plasma_params;
int ckd_plasma_draw(ckd_ctx *ctx, const ckd_plasma_params *p, float time, uint32_t *d_dest); /* Plasma_Draw shadertoy.cpp:276 */

typedef struct ckd_nautilus_params {     /* shadertoy.cpp:304-307,399 */
	float roll;          /* nautilus:Roll */
	float hue;           /* nautilus:Hue */
	float speed;         /* nautilus:Speed */
	float desaturation;  /* nautilus:Desat */
	float blur;          /* nautilus:Blur (raw track value, BoxBlurScale is applied inside) */
} ckd_nautilus_params;
int ckd_nautilus_draw(ckd_ctx *ctx, const ckd_nautilus_params *p, float time, uint32_t *d_dest); /* Nautilus_Draw shadertoy.cpp:395 */

typedef struct ckd_spikey_params {       /* shadertoy.cpp:436-450, 529-537, 604-606, 669-677, 715 */
	float speed;             /* spike:Speed */
	float roll;              /* spike:Roll */
	float specular;          /* spike:Specular */
	float desaturation;      /* spike:Desaturation */
	float hue;               /* spike:Hue */
	float gamma;             /* spike:Gamma */
	float warmup;            /* distSpike:Warmup */
	float dist_x, dist_y, dist_z;    /* distSpike:xOffs/yOffs/zOffs */
	float close_x, close_y, close_z; /* closeSpike:xOffs/yOffs/zOffs */
	float close_z_scale;     /* closeSpike:zOffsScale */
	float close_normal_grain;/* closeSpike:NormalGrain */
	float close_scale;       /* closeSpike:Scale */
	int close_rim;           /* geti(closeSpike:Rim) */
	int close_aspect_mul;    /* geti(closeSpike:AspectMul) */
	float mix_blur_map;      /* closeSpike:MixBlurMap */
	float mix_blur;          /* closeSpike:MixBlur */
	float mix_map_blur;      /* closeSpike:MixMapBlur */
	float mix_blur_opacity;  /* closeSpike:MixBlurOpacity */
} ckd_spikey_params;
int ckd_spikey_draw(ckd_ctx *ctx, const ckd_spikey_params *p, float time, int close, uint32_t *d_dest); /* Spikey_Draw shadertoy.cpp:661 */

typedef struct ckd_tunnel_params {       /* shadertoy.cpp:752-762, 842-843 */
	float boxy, flower_scale, flower_freq, flower_phase; /* tunnel:Boxy/FlowerScale/FlowerFreq/FlowerPhase */
	float speed, roll, pitch, radius;                    /* tunnel:Speed/Roll/Pitch/Radius */
	float mul_u, mul_v;                                  /* tunnel:MulU/MulV */
	int lit_tiles;                                       /* geti(tunnel:LitTiles) */
	float lit_blur;                                      /* tunnel:LitBlur */
	float fog1, fog2;                                    /* tunnel:Fog1/Fog2 */
} ckd_tunnel_params;
int ckd_tunnel_draw(ckd_ctx *ctx, const ckd_tunnel_params *p, float time, uint32_t *d_dest); /* Tunnel_Draw shadertoy.cpp:840 */

typedef struct ckd_sinuses_params {      /* shadertoy.cpp:905-911 */
	float specular, roll, speed, offs_x, gamma, hue, desaturation; /* sinusesTunnel:* */
} ckd_sinuses_params;
int ckd_sinuses_draw(ckd_ctx *ctx, const ckd_sinuses_params *p, float time, uint32_t *d_dest); /* Sinuses_Draw shadertoy.cpp:984 */

typedef struct ckd_laura_params {        /* shadertoy.cpp:1019-1024 */
	float speed, yaw, pitch, roll, hue, saturate; /* laura:* */
} ckd_laura_params;
int ckd_laura_draw(ckd_ctx *ctx, const ckd_laura_params *p, float time, uint32_t *d_dest); /* Laura_Draw shadertoy.cpp:1105 */

typedef struct ckd_landscape_params {    /* landscape.cpp:128,156,230-242 (gamepad state = 0: view angle 0, no strafe) */
	float forward;        /* voxelScape:Forward */
	float tilt;           /* voxelScape:Tilt */
	float warp_speed;     /* voxelScape:WarpSpeed */
	float warp_strength;  /* voxelScape:WarpStrength */
	/* accumulated gamepad state (landscape.cpp:116-154 statics); all zero for a headless run */
	float pad_tilt, view_angle, strafe_x, strafe_y, pad_move_x, pad_move_y;
} ckd_landscape_params;
int ckd_landscape_draw(ckd_ctx *ctx, const ckd_landscape_params *p, float time, uint32_t *d_dest); /* Landscape_Draw landscape.cpp:228 */

typedef struct ckd_tunnelscape_params {  /* tunnelscape.cpp:108-111,177 */
	float step_u, step_v, speed, blur; /* starsTunnel:stepU/stepV/Speed/Blur */
} ckd_tunnelscape_params;
int ckd_tunnelscape_draw(ckd_ctx *ctx, const ckd_tunnelscape_params *p, float time, uint32_t *d_dest); /* Tunnelscape_Draw tunnelscape.cpp:168 */

typedef struct ckd_ball_params {         /* ball.cpp:82,285-286,318-319,322,331-332,456-470,483,486 */
	float blur;             /* ball:Blur */
	float radius;           /* ball:Radius */
	int ray_length;         /* geti(ball:RayLength) */
	int spikes;             /* geti(ball:Spikes) */
	int has_beams;          /* geti(ball:HasBeams) */
	int base_shape_index;   /* geti(ball:BaseShapeIndex) */
	float speed;            /* ball:Speed */
	int beam_atten;         /* geti(ball:BeamAttenuation) */
	float beam_alpha_min;   /* ball:BeamAlphaMin */
	float rotate_offs_x, rotate_offs_y; /* ball:RotateOffsX/Y */
	float beams1, beams2, beams3;       /* ball:Beams1..3 */
	int low_beams;          /* geti(ball:BallLowBeams) */
} ckd_ball_params;
int ckd_ball_draw(ckd_ctx *ctx, const ckd_ball_params *p, float time, uint32_t *d_dest); /* Ball_Draw ball.cpp:452 */
/* The beam tail of vball_ray_beams on its own (ball.cpp:168-203: beam colour with a smoothstep alpha whose parameter is a float
 * accumulated pixel by pixel), for `rows` rows of `row_pixels` pixels in device memory: row r is a ray whose spans ended
 * `first_remainder + r` pixels before the last pixel of its row.  Pixels the tail does not reach are left alone.  The ball
 * kernel runs the same code per ray; this entry exists so that every tail length can be checked on its own: with raw_steps != 0
 * it stores the float bits of the accumulated parameter (`curStep`, ball.cpp:195-203) instead of the pixels. */
int ckd_ball_beam_tail(ckd_ctx *ctx, uint32_t *d_rows, int row_pixels, int rows, int first_remainder, uint32_t beam_color, float beam_alpha_min, int raw_steps);

typedef struct ckd_twister_params {      /* torus-twister.cpp:100-101,173 */
	float speed, shear_speed, blur; /* twister:Speed, twister::ShearSpeed (sic), twister:Blur */
} ckd_twister_params;
int ckd_twister_draw(ckd_ctx *ctx, const ckd_twister_params *p, float time, uint32_t *d_dest); /* Twister_Draw torus-twister.cpp:166 */

/* ---- frame gather over peer memory (frames shard over the GPUs of a box, frame i -> rank i mod N, one process
 *      per GPU; the finished frames meet on one GPU before they go where the reference hands them to Display::Update,
 *      display.cpp:66-82).  A ring of frame slots in the collector's HBM, written by the producers with peer copies over
 *      NVLink (CUDA IPC mapping, copy engines) and flag-signalled on the device: no NCCL, no host round trip per frame.
 *      Frames carry a global sequence number and are pushed / popped in the order 0, 1, 2, ...                          ---- */
typedef struct ckd_gather ckd_gather;
#define CKD_GATHER_HANDLE_BYTES 128
#define CKD_GATHER_CHECKSUM 1   /* ckd_gather_pop: fold the frame into the checksum table (sum_i pixel[i]*(2i+1) mod 2^64) */
#define CKD_GATHER_TO_HOST  2   /* ckd_gather_pop: copy the frame to h_dest */
/* collector (the process whose GPU receives the frames): allocates `slots` (2..64) frame slots of the context's resolution */
int ckd_gather_create(ckd_ctx *ctx, int slots, ckd_gather **out_gather);
/* collector: writes CKD_GATHER_HANDLE_BYTES bytes that another process hands to ckd_gather_open (over any channel) */
int ckd_gather_export(ckd_gather *gather, void *out_handle);
/* producer in another process (same or another GPU of the box): maps the collector's ring */
int ckd_gather_open(ckd_ctx *ctx, const void *handle, ckd_gather **out_gather);
void ckd_gather_destroy(ckd_gather *gather);
int ckd_gather_set_timeout_ms(ckd_gather *gather, unsigned timeout_ms); /* device-side waits give up after this long (default 20 s) */
/* producer: a local device frame to render the next frame into (three are handed out in turn; the context's stream waits,
 * on the device, for the copy that last read it) */
int ckd_gather_acquire(ckd_gather *gather, uint32_t **out_d_frame);
/* producer: publish a frame as sequence number seq, ordered after everything enqueued on the context's stream so far.
 * d_frame == NULL: the frame of the last ckd_gather_acquire; otherwise any device frame that stays untouched until
 * ckd_gather_flush + ckd_sync.  Waits on the device (not the host) until the collector has drained slot seq % slots. */
int ckd_gather_push(ckd_gather *gather, const uint32_t *d_frame, unsigned long long seq);
/* the same for a frame rendered on another context of this process and device (ckd_own_stream / ckd_clone_inputs): the staging
 * frame is handed out against, and the push ordered after, `render_ctx`'s stream.  Eight staging frames are handed out in turn. */
int ckd_gather_acquire_on(ckd_gather *gather, ckd_ctx *render_ctx, uint32_t **out_d_frame);
int ckd_gather_push_on(ckd_gather *gather, ckd_ctx *render_ctx, const uint32_t *d_frame, unsigned long long seq);
/* collector: consume sequence number seq (call in order).  mode: 0 release the slot, CKD_GATHER_CHECKSUM, CKD_GATHER_TO_HOST
 * (h_dest page-locked), or both.  Enqueued on the gather's consumer stream. */
int ckd_gather_pop(ckd_gather *gather, unsigned long long seq, int mode, void *h_dest);
int ckd_gather_wait_pop(ckd_gather *gather, unsigned long long seq); /* host blocks until that pop is done (one of the last 16) */
/* the context's stream waits for everything pushed / popped so far (so an event recorded on it afterwards covers the gather) */
int ckd_gather_flush(ckd_gather *gather);
int ckd_gather_status(ckd_gather *gather); /* synchronises; CKD_OK or CKD_ERR_TIMEOUT */
int ckd_gather_checksums(ckd_gather *gather, unsigned long long first_seq, unsigned count, unsigned long long *out_sums);
unsigned long long ckd_gather_peer_bytes(const ckd_gather *gather); /* bytes this process has copied into the collector's memory */
int ckd_gather_slots(const ckd_gather *gather);
/* the same checksum for a frame in this context's memory (synchronises) */
int ckd_frame_checksum(ckd_ctx *ctx, const uint32_t *d_frame, unsigned long long *out_sum);

/* number of CUDA kernels this library launched on this context so far (bench.py's gpu_launches) */
unsigned long long ckd_launch_count(const ckd_ctx *ctx);

/* per-kernel timing with CUDA events on the context's stream (measurement only; replaces nothing in the reference,
 * whose only meter is the average-FPS counter of main.cpp:352-356).  ckd_profile_begin() starts recording a pair of
 * events around every kernel launch; ckd_profile_end() synchronises and returns one aggregate per kernel name. */
typedef struct ckd_kernel_stat {
	char name[48];
	unsigned launches;
	double total_ms;       /* sum of the per-launch event durations */
	double algo_bytes;     /* sum of the algorithmic bytes of those launches (SURVEY.md 8d / DESIGN.md) */
} ckd_kernel_stat;
int ckd_profile_begin(ckd_ctx *ctx);
int ckd_profile_end(ckd_ctx *ctx, ckd_kernel_stat *out_stats, int max_stats, int *out_count);

#ifdef __cplusplus
}
#endif

#endif /* CKD_H_ */
