// ckd_image.cpp -- the asset decode path of the host layer: Image_Create/Image_Load32/Image_Load8/
// Image_Load32_CA of the reference (image.cpp:13-110, image.h:7-17) without DevIL.
//
// The reference hands every file to DevIL and asks for BGRA bytes (= little-endian 0xAARRGGBB) or 8-bit luminance with
// the origin in the upper left corner (image.cpp:16-17, 49-60).  Its art is 90 PNG files (8-bit grey / RGB / RGBA /
// palette with tRNS, not interlaced) and 15 JPEG files (8-bit YCbCr, all components sampled 1x1, twelve baseline and
// three progressive); both decoders are written out here:
//   * PNG (ISO/IEC 15948): chunk walk with CRC check, zlib inflate (libz, the one library DevIL's libpng uses as well),
//     the five scanline filters, every colour type and bit depth, Adam7.  Lossless, so the pixels are the file's pixels.
//   * JPEG (ITU-T T.81): Huffman baseline and progressive (spectral selection + successive approximation), restart
//     intervals, the IJG "islow" inverse DCT (Loeffler-Ligtenberg-Moschytz, 13-bit constants, two passes) and the IJG
//     fixed-point YCbCr->RGB tables, triangle ("fancy") chroma upsampling for 2x1 and 2x2 subsampled files -- the
//     decoder libjpeg runs by default, which is what DevIL and Pillow both link, so the bytes agree with the pixels the
//     test harness shares between the reference and this library (refdata/assets.npz; tests/test_image_decode.py).
// Conversions: grey -> BGRA replicates the value; palette entries take their tRNS alpha; a tRNS colour key clears alpha;
// 16-bit samples keep their high byte (png_set_strip_16, as DevIL does); colour -> luminance for Image_Load8 uses the
// ITU-R 601 weights in 16.16 fixed point (the harness' rule; exact for the grey art the reference loads that way).
// PNG gAMA is ignored (DevIL applies it only as screen 2.2 x file 0.45455 = 1.0).

#include "ckd_host_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <zlib.h>

namespace {

std::string s_assetRoot;          // prefix for relative paths; the reference runs with cwd = target/
std::vector<void *> s_gc;         // s_pGC, image.cpp:11

struct Decoded
{
	int width = 0, height = 0;
	int channels = 0;             // 1 = L, 2 = LA, 3 = RGB, 4 = RGBA (8 bits per sample, row-major, top row first)
	std::vector<uint8_t> px;
};

bool Fail(const std::string &path, const char *why)
{
	SetLastError("Can not load image: " + path + " (" + why + ")"); // image.cpp:40
	return false;
}

bool ReadFile(const std::string &path, std::vector<uint8_t> &bytes)
{
	const std::string full = (!s_assetRoot.empty() && !path.empty() && path[0] != '/') ? s_assetRoot + "/" + path : path;
	FILE *fp = fopen(full.c_str(), "rb");
	if (!fp)
		return false;
	fseek(fp, 0, SEEK_END);
	const long size = ftell(fp);
	fseek(fp, 0, SEEK_SET);
	bytes.resize(size > 0 ? size_t(size) : 0);
	const bool ok = bytes.empty() || 1 == fread(bytes.data(), bytes.size(), 1, fp);
	fclose(fp);
	return ok && !bytes.empty();
}

inline uint32_t Be32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
inline unsigned Be16(const uint8_t *p) { return (unsigned(p[0]) << 8) | p[1]; }

// ---------------------------------------------------------------------------------------------------------------
// PNG
// ---------------------------------------------------------------------------------------------------------------

inline int Paeth(int a, int b, int c)
{
	const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
	return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// undo the scanline filters of one (sub)image in place; `raw` holds height x (1 + rowBytes) bytes
bool Unfilter(uint8_t *raw, size_t rowBytes, int height, int bpp /* bytes per complete pixel, >= 1 */)
{
	std::vector<uint8_t> zero(rowBytes, 0);
	const uint8_t *prev = zero.data();
	for (int y = 0; y < height; ++y)
	{
		uint8_t *line = raw + size_t(y)*(rowBytes + 1);
		const int filter = line[0];
		uint8_t *cur = line + 1;
		switch (filter)
		{
		case 0: break;
		case 1: for (size_t i = bpp; i < rowBytes; ++i) cur[i] = uint8_t(cur[i] + cur[i - bpp]); break;
		case 2: for (size_t i = 0; i < rowBytes; ++i) cur[i] = uint8_t(cur[i] + prev[i]); break;
		case 3:
			for (size_t i = 0; i < rowBytes; ++i)
				cur[i] = uint8_t(cur[i] + (((i >= size_t(bpp) ? cur[i - bpp] : 0) + prev[i]) >> 1));
			break;
		case 4:
			for (size_t i = 0; i < rowBytes; ++i)
				cur[i] = uint8_t(cur[i] + Paeth(i >= size_t(bpp) ? cur[i - bpp] : 0, prev[i], i >= size_t(bpp) ? prev[i - bpp] : 0));
			break;
		default: return false;
		}
		prev = cur;
	}
	return true;
}

bool DecodePng(const std::string &path, const std::vector<uint8_t> &file, Decoded &out)
{
	static const uint8_t kSignature[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
	if (file.size() < 8 + 25 || 0 != memcmp(file.data(), kSignature, 8))
		return Fail(path, "not a PNG file");

	uint32_t width = 0, height = 0;
	int depth = 0, colorType = -1, interlace = 0;
	std::vector<uint8_t> idat, palette, trns;
	bool haveHeader = false, sawEnd = false;
	for (size_t pos = 8; pos + 12 <= file.size() && !sawEnd; )
	{
		const uint32_t length = Be32(&file[pos]);
		const uint8_t *type = &file[pos + 4], *data = &file[pos + 8];
		if (length > file.size() - pos - 12)
			return Fail(path, "truncated PNG chunk");
		if (uint32_t(crc32(crc32(0L, Z_NULL, 0), type, length + 4)) != Be32(data + length))
			return Fail(path, "PNG chunk CRC mismatch");
		if (0 == memcmp(type, "IHDR", 4) && 13 == length)
		{
			width = Be32(data); height = Be32(data + 4);
			depth = data[8]; colorType = data[9]; interlace = data[12];
			if (0 != data[10] || 0 != data[11] || interlace > 1)
				return Fail(path, "unknown PNG compression, filter or interlace method");
			haveHeader = true;
		}
		else if (0 == memcmp(type, "PLTE", 4)) palette.assign(data, data + length);
		else if (0 == memcmp(type, "tRNS", 4)) trns.assign(data, data + length);
		else if (0 == memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + length);
		else if (0 == memcmp(type, "IEND", 4)) sawEnd = true;
		pos += size_t(length) + 12;
	}
	static const int kSamples[7] = { 1, 0, 3, 1, 2, 0, 4 };
	if (!haveHeader || colorType < 0 || colorType > 6 || 0 == kSamples[colorType] || 0 == width || 0 == height || width > 16384 || height > 16384)
		return Fail(path, "bad PNG header");
	if (!(depth == 8 || depth == 16 || ((colorType == 0 || colorType == 3) && (depth == 1 || depth == 2 || depth == 4))) || (colorType == 3 && depth == 16))
		return Fail(path, "bad PNG bit depth");
	if (3 == colorType && palette.size() < 3)
		return Fail(path, "PNG palette missing");

	const int samples = kSamples[colorType];
	const int bitsPerPixel = samples*depth;
	const int bpp = bitsPerPixel >= 8 ? bitsPerPixel/8 : 1;
	auto rowBytesOf = [&](uint32_t w) { return (size_t(w)*bitsPerPixel + 7)/8; };

	// Adam7 (interlace 1): seven reduced images, each filtered on its own; interlace 0 is the single full pass
	static const int kX0[7] = { 0, 4, 0, 2, 0, 1, 0 }, kY0[7] = { 0, 0, 4, 0, 2, 0, 1 }, kDX[7] = { 8, 8, 4, 4, 2, 2, 1 }, kDY[7] = { 8, 8, 8, 4, 4, 2, 2 };
	struct Pass { uint32_t w, h; int x0, y0, dx, dy; size_t offset; };
	std::vector<Pass> passes;
	size_t rawSize = 0;
	if (0 == interlace)
	{
		passes.push_back({ width, height, 0, 0, 1, 1, 0 });
		rawSize = size_t(height)*(rowBytesOf(width) + 1);
	}
	else
		for (int i = 0; i < 7; ++i)
		{
			const uint32_t w = (width + kDX[i] - 1 - kX0[i])/kDX[i], h = (height + kDY[i] - 1 - kY0[i])/kDY[i];
			if (0 == w || 0 == h) continue;
			passes.push_back({ w, h, kX0[i], kY0[i], kDX[i], kDY[i], rawSize });
			rawSize += size_t(h)*(rowBytesOf(w) + 1);
		}

	std::vector<uint8_t> raw(rawSize);
	uLongf got = uLongf(rawSize);
	const int zrc = uncompress(raw.data(), &got, idat.data(), uLong(idat.size()));
	if ((Z_OK != zrc && Z_BUF_ERROR != zrc) || got != rawSize)
		return Fail(path, "PNG pixel data does not inflate to the image size");

	// tRNS: per-entry alpha for palettes, a colour key for grey / RGB
	const bool keyed = !trns.empty() && (0 == colorType || 2 == colorType);
	const bool paletteAlpha = 3 == colorType && !trns.empty();
	unsigned key[3] = { 0, 0, 0 };
	if (keyed)
	{
		if (trns.size() < size_t(0 == colorType ? 2 : 6)) return Fail(path, "bad PNG tRNS chunk");
		for (int c = 0; c < (0 == colorType ? 1 : 3); ++c) key[c] = Be16(&trns[c*2]);
	}

	const bool hasAlpha = 4 == colorType || 6 == colorType || keyed || paletteAlpha;
	const bool isColor = 2 == colorType || 3 == colorType || 6 == colorType;
	out.width = int(width); out.height = int(height);
	out.channels = (isColor ? 3 : 1) + (hasAlpha ? 1 : 0);
	out.px.assign(size_t(width)*height*out.channels, 0);

	for (const Pass &pass : passes)
	{
		const size_t rowBytes = rowBytesOf(pass.w);
		uint8_t *base = raw.data() + pass.offset;
		if (!Unfilter(base, rowBytes, int(pass.h), bpp))
			return Fail(path, "unknown PNG scanline filter");
		for (uint32_t py = 0; py < pass.h; ++py)
		{
			const uint8_t *line = base + size_t(py)*(rowBytes + 1) + 1;
			const uint32_t y = pass.y0 + py*pass.dy;
			for (uint32_t px = 0; px < pass.w; ++px)
			{
				unsigned s[4] = { 0, 0, 0, 0 };     // samples at file precision
				if (depth < 8)
				{
					const size_t bit = size_t(px)*depth;
					s[0] = (line[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1);
				}
				else if (8 == depth)
					for (int c = 0; c < samples; ++c) s[c] = line[size_t(px)*samples + c];
				else
					for (int c = 0; c < samples; ++c) s[c] = Be16(&line[(size_t(px)*samples + c)*2]);

				uint8_t *dst = &out.px[(size_t(y)*width + pass.x0 + size_t(px)*pass.dx)*out.channels];
				auto to8 = [&](unsigned v) -> uint8_t
				{
					if (16 == depth) return uint8_t(v >> 8);           // png_set_strip_16
					if (8 == depth) return uint8_t(v);
					return uint8_t(v*255u/((1u << depth) - 1));         // 1/2/4-bit grey expands to the full range
				};
				switch (colorType)
				{
				case 0:
					dst[0] = to8(s[0]);
					if (keyed) dst[1] = (s[0] == key[0]) ? 0 : 255;
					break;
				case 2:
					dst[0] = to8(s[0]); dst[1] = to8(s[1]); dst[2] = to8(s[2]);
					if (keyed) dst[3] = (s[0] == key[0] && s[1] == key[1] && s[2] == key[2]) ? 0 : 255;
					break;
				case 3:
					if (size_t(s[0])*3 + 2 < palette.size()) { dst[0] = palette[s[0]*3]; dst[1] = palette[s[0]*3 + 1]; dst[2] = palette[s[0]*3 + 2]; }
					if (paletteAlpha) dst[3] = s[0] < trns.size() ? trns[s[0]] : 255;
					break;
				case 4: dst[0] = to8(s[0]); dst[1] = to8(s[1]); break;
				case 6: dst[0] = to8(s[0]); dst[1] = to8(s[1]); dst[2] = to8(s[2]); dst[3] = to8(s[3]); break;
				}
			}
		}
	}
	return true;
}

// ---------------------------------------------------------------------------------------------------------------
// JPEG (ITU-T T.81), baseline and progressive Huffman, 8 bits per sample
// ---------------------------------------------------------------------------------------------------------------

const uint8_t kZigZag[64 + 16] = {
	 0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28,
	35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
	63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63 }; // the tail keeps corrupt run lengths inside the block

struct HuffTable
{
	bool present = false;
	uint8_t bits[17] = {}, values[256] = {};
	int minCode[17], maxCode[18], valPtr[17];

	void Build()
	{
		int code = 0, k = 0;
		for (int len = 1; len <= 16; ++len)
		{
			valPtr[len] = k;
			minCode[len] = code;
			code += bits[len];
			k += bits[len];
			maxCode[len] = bits[len] ? code - 1 : -1;
			code <<= 1;
		}
		maxCode[17] = 0x7fffffff;
		present = true;
	}
};

struct Component
{
	int id = 0, h = 1, v = 1, tq = 0;
	int dcTable = 0, acTable = 0;
	int blocksW = 0, blocksH = 0;     // allocated blocks (padded to whole MCUs)
	int widthInBlocks = 0, heightInBlocks = 0; // blocks that carry image samples (non-interleaved scans stop here)
	int dcPred = 0;
	std::vector<int16_t> coef;        // blocksW*blocksH*64, natural order
	std::vector<uint8_t> plane;       // blocksW*8 x blocksH*8 samples after the inverse DCT
};

struct BitReader
{
	const uint8_t *p, *end;
	uint32_t acc = 0;
	int count = 0;
	bool hitMarker = false;

	void Fill()
	{
		while (count <= 24)
		{
			unsigned byte = 0;
			if (!hitMarker && p < end)
			{
				byte = *p;
				if (0xff == byte)
				{
					if (p + 1 < end && 0 == p[1]) p += 2;      // stuffed zero
					else { hitMarker = true; byte = 0; }       // a marker ends the entropy-coded segment: feed zeros
				}
				else
					++p;
			}
			acc |= byte << (24 - count);
			count += 8;
		}
	}
	int Bit() { if (count < 1) Fill(); const int b = int(acc >> 31); acc <<= 1; --count; return b; }
	int Bits(int n) { if (0 == n) return 0; if (count < n) Fill(); const int v = int(acc >> (32 - n)); acc <<= n; count -= n; return v; }
	void Reset() { acc = 0; count = 0; hitMarker = false; }
};

inline int Extend(int v, int n) { return (n && v < (1 << (n - 1))) ? v - (1 << n) + 1 : v; } // T.81 F.12

struct JpegDecoder
{
	const std::string &path;
	const std::vector<uint8_t> &file;
	uint16_t quant[4][64] = {};
	HuffTable dc[4], ac[4];
	std::vector<Component> comps;
	int width = 0, height = 0, hMax = 1, vMax = 1, mcusX = 0, mcusY = 0;
	bool progressive = false, haveFrame = false;
	int restartInterval = 0;
	int adobeTransform = -1;
	BitReader br{ nullptr, nullptr };
	int eobRun = 0;

	JpegDecoder(const std::string &path_, const std::vector<uint8_t> &file_) : path(path_), file(file_) {}

	int DecodeSymbol(const HuffTable &t)
	{
		int code = br.Bit();
		int len = 1;
		while (len <= 16 && code > t.maxCode[len])
		{
			code = (code << 1) | br.Bit();
			++len;
		}
		if (len > 16) return 0;
		return t.values[(t.valPtr[len] + code - t.minCode[len]) & 255];
	}

	// --- block decoders ---------------------------------------------------------------------------------------
	void BaselineBlock(Component &c, int16_t *blk)
	{
		const int t = DecodeSymbol(dc[c.dcTable]) & 15; // categories 0..11 (a corrupt table cannot ask for more bits than the reader holds)
		c.dcPred += Extend(br.Bits(t), t);
		blk[0] = int16_t(c.dcPred);
		for (int k = 1; k < 64; )
		{
			const int rs = DecodeSymbol(ac[c.acTable]), r = rs >> 4, s = rs & 15;
			if (0 == s)
			{
				if (15 != r) break;
				k += 16;
				continue;
			}
			k += r;
			blk[kZigZag[k]] = int16_t(Extend(br.Bits(s), s));
			++k;
		}
	}

	void DcFirst(Component &c, int16_t *blk, int al)
	{
		const int t = DecodeSymbol(dc[c.dcTable]) & 15; // categories 0..11 (a corrupt table cannot ask for more bits than the reader holds)
		c.dcPred += Extend(br.Bits(t), t);
		blk[0] = int16_t(c.dcPred*(1 << al));
	}

	void DcRefine(int16_t *blk, int al) { if (br.Bit()) blk[0] |= int16_t(1 << al); }

	void AcFirst(Component &c, int16_t *blk, int ss, int se, int al)
	{
		if (eobRun > 0) { --eobRun; return; }
		for (int k = ss; k <= se; )
		{
			const int rs = DecodeSymbol(ac[c.acTable]), r = rs >> 4, s = rs & 15;
			if (0 == s)
			{
				if (r < 15)
				{
					eobRun = (1 << r) - 1;
					if (r) eobRun += br.Bits(r);
					break;
				}
				k += 16;
				continue;
			}
			k += r;
			blk[kZigZag[k]] = int16_t(Extend(br.Bits(s), s)*(1 << al));
			++k;
		}
	}

	void AcRefine(Component &c, int16_t *blk, int ss, int se, int al)
	{
		const int p1 = 1 << al, m1 = -1*(1 << al);
		int k = ss;
		if (eobRun <= 0)
		{
			for (; k <= se; ++k)
			{
				const int rs = DecodeSymbol(ac[c.acTable]);
				int r = rs >> 4;
				const int s = rs & 15;
				int value = 0;
				if (s)
					value = br.Bit() ? p1 : m1;      // a newly non-zero coefficient (always magnitude 1 at this bit)
				else if (15 != r)
				{
					eobRun = 1 << r;
					if (r) eobRun += br.Bits(r);
					break;
				}
				// skip r still-zero coefficients, refining the already non-zero ones on the way
				for (; k <= se; ++k)
				{
					int16_t &coef = blk[kZigZag[k]];
					if (coef)
					{
						if (br.Bit() && 0 == (coef & p1))
							coef = int16_t(coef >= 0 ? coef + p1 : coef + m1);
					}
					else if (--r < 0)
						break;
				}
				if (value && k <= se)
					blk[kZigZag[k]] = int16_t(value);
			}
		}
		if (eobRun > 0)
		{
			for (; k <= se; ++k)
			{
				int16_t &coef = blk[kZigZag[k]];
				if (coef && br.Bit() && 0 == (coef & p1))
					coef = int16_t(coef >= 0 ? coef + p1 : coef + m1);
			}
			--eobRun;
		}
	}

	// --- one scan ---------------------------------------------------------------------------------------------
	bool Scan(size_t &pos, const std::vector<int> &scanComps, int ss, int se, int ah, int al)
	{
		br.p = &file[pos]; br.end = file.data() + file.size();
		br.Reset();
		eobRun = 0;
		for (int ci : scanComps) comps[ci].dcPred = 0;

		const bool interleaved = scanComps.size() > 1;
		Component &first = comps[scanComps[0]];
		const int unitsX = interleaved ? mcusX : first.widthInBlocks, unitsY = interleaved ? mcusY : first.heightInBlocks;
		int untilRestart = restartInterval;

		auto decodeBlock = [&](Component &c, int bx, int by)
		{
			int16_t *blk = &c.coef[(size_t(by)*c.blocksW + bx)*64];
			if (!progressive) BaselineBlock(c, blk);
			else if (0 == ss) { if (0 == ah) DcFirst(c, blk, al); else DcRefine(blk, al); }
			else if (0 == ah) AcFirst(c, blk, ss, se, al);
			else AcRefine(c, blk, ss, se, al);
		};

		for (int uy = 0; uy < unitsY; ++uy)
			for (int ux = 0; ux < unitsX; ++ux)
			{
				if (restartInterval && 0 == untilRestart)
				{
					// RSTn: byte align, skip the marker, reset the predictors
					const uint8_t *q = br.p;
					while (q + 1 < br.end && !(0xff == q[0] && q[1] >= 0xd0 && q[1] <= 0xd7)) ++q;
					if (q + 1 >= br.end) return Fail(path, "JPEG restart marker missing");
					br.p = q + 2;
					br.Reset();
					eobRun = 0;
					for (int ci : scanComps) comps[ci].dcPred = 0;
					untilRestart = restartInterval;
				}
				if (interleaved)
				{
					for (int ci : scanComps)
					{
						Component &c = comps[ci];
						for (int v = 0; v < c.v; ++v)
							for (int h = 0; h < c.h; ++h)
								decodeBlock(c, ux*c.h + h, uy*c.v + v);
					}
				}
				else
					decodeBlock(first, ux, uy);
				--untilRestart;
			}

		// continue the marker walk behind the entropy-coded data
		const uint8_t *q = br.p;
		while (q + 1 < br.end && !(0xff == q[0] && 0 != q[1] && !(q[1] >= 0xd0 && q[1] <= 0xd7) && 0xff != q[1])) ++q;
		pos = size_t(q - file.data());
		return true;
	}

	// --- IJG jidctint.c "islow": LL&M, CONST_BITS 13, PASS1_BITS 2 --------------------------------------------------
	static inline uint8_t Clamp(long v) { v = (v >> 0) + 128; return uint8_t(v < 0 ? 0 : (v > 255 ? 255 : v)); }

	static void Idct(const int16_t *coef, const uint16_t *q, uint8_t *out, size_t stride)
	{
		const long F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299, F1_847 = 15137, F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
		const int CONST_BITS = 13, PASS1_BITS = 2;
		auto descale = [](long x, int n) { return (x + (1L << (n - 1))) >> n; };
		long ws[64];
		for (int c = 0; c < 8; ++c)
		{
			const int16_t *in = coef + c;
			const uint16_t *qt = q + c;
			long *w = ws + c;
			if (0 == (in[8] | in[16] | in[24] | in[32] | in[40] | in[48] | in[56]))
			{
				const long dcval = long(in[0])*qt[0]*(1L << PASS1_BITS);
				for (int r = 0; r < 8; ++r) w[r*8] = dcval;
				continue;
			}
			long z2 = long(in[16])*qt[16], z3 = long(in[48])*qt[48];
			long z1 = (z2 + z3)*F0_541;
			long tmp2 = z1 + z3*(-F1_847), tmp3 = z1 + z2*F0_765;
			z2 = long(in[0])*qt[0]; z3 = long(in[32])*qt[32];
			long tmp0 = (z2 + z3)*(1L << CONST_BITS), tmp1 = (z2 - z3)*(1L << CONST_BITS);
			const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
			tmp0 = long(in[56])*qt[56]; tmp1 = long(in[40])*qt[40]; tmp2 = long(in[24])*qt[24]; tmp3 = long(in[8])*qt[8];
			z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
			long z4 = tmp1 + tmp3;
			const long z5 = (z3 + z4)*F1_175;
			tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
			z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
			z3 += z5; z4 += z5;
			tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
			w[0]  = descale(tmp10 + tmp3, CONST_BITS - PASS1_BITS); w[56] = descale(tmp10 - tmp3, CONST_BITS - PASS1_BITS);
			w[8]  = descale(tmp11 + tmp2, CONST_BITS - PASS1_BITS); w[48] = descale(tmp11 - tmp2, CONST_BITS - PASS1_BITS);
			w[16] = descale(tmp12 + tmp1, CONST_BITS - PASS1_BITS); w[40] = descale(tmp12 - tmp1, CONST_BITS - PASS1_BITS);
			w[24] = descale(tmp13 + tmp0, CONST_BITS - PASS1_BITS); w[32] = descale(tmp13 - tmp0, CONST_BITS - PASS1_BITS);
		}
		for (int r = 0; r < 8; ++r)
		{
			const long *w = ws + r*8;
			uint8_t *o = out + size_t(r)*stride;
			if (0 == (w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]))
			{
				const uint8_t dcval = Clamp(descale(w[0], PASS1_BITS + 3));
				for (int c = 0; c < 8; ++c) o[c] = dcval;
				continue;
			}
			long z2 = w[2], z3 = w[6];
			long z1 = (z2 + z3)*F0_541;
			long tmp2 = z1 + z3*(-F1_847), tmp3 = z1 + z2*F0_765;
			long tmp0 = (w[0] + w[4])*(1L << CONST_BITS), tmp1 = (w[0] - w[4])*(1L << CONST_BITS);
			const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
			tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
			z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
			long z4 = tmp1 + tmp3;
			const long z5 = (z3 + z4)*F1_175;
			tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
			z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
			z3 += z5; z4 += z5;
			tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
			const int S = CONST_BITS + PASS1_BITS + 3;
			o[0] = Clamp(descale(tmp10 + tmp3, S)); o[7] = Clamp(descale(tmp10 - tmp3, S));
			o[1] = Clamp(descale(tmp11 + tmp2, S)); o[6] = Clamp(descale(tmp11 - tmp2, S));
			o[2] = Clamp(descale(tmp12 + tmp1, S)); o[5] = Clamp(descale(tmp12 - tmp1, S));
			o[3] = Clamp(descale(tmp13 + tmp0, S)); o[4] = Clamp(descale(tmp13 - tmp0, S));
		}
	}

	// --- marker walk ------------------------------------------------------------------------------------------
	bool Decode(Decoded &out)
	{
		if (file.size() < 4 || 0xff != file[0] || 0xd8 != file[1])
			return Fail(path, "not a JPEG file");
		size_t pos = 2;
		bool done = false;
		while (!done && pos + 4 <= file.size())
		{
			if (0xff != file[pos]) { ++pos; continue; }
			const int marker = file[pos + 1];
			if (0xff == marker) { ++pos; continue; }
			if (0xd9 == marker) break;
			if (0x01 == marker || (marker >= 0xd0 && marker <= 0xd7)) { pos += 2; continue; }
			const size_t length = Be16(&file[pos + 2]);
			if (length < 2 || pos + 2 + length > file.size()) return Fail(path, "truncated JPEG segment");
			const uint8_t *seg = &file[pos + 4], *segEnd = &file[pos + 2 + length];
			pos += 2 + length;
			switch (marker)
			{
			case 0xdb: // DQT
				while (seg < segEnd)
				{
					const int pq = seg[0] >> 4, tq = seg[0] & 15;
					++seg;
					if (tq > 3 || seg + (pq ? 128 : 64) > segEnd) return Fail(path, "bad JPEG quantisation table");
					for (int i = 0; i < 64; ++i, seg += pq ? 2 : 1)
						quant[tq][kZigZag[i]] = uint16_t(pq ? Be16(seg) : seg[0]);
				}
				break;
			case 0xc4: // DHT
				while (seg + 17 <= segEnd)
				{
					const int tc = seg[0] >> 4, th = seg[0] & 15;
					if (tc > 1 || th > 3) return Fail(path, "bad JPEG Huffman table");
					HuffTable &t = tc ? ac[th] : dc[th];
					int total = 0;
					t.bits[0] = 0;
					for (int i = 1; i <= 16; ++i) total += (t.bits[i] = seg[i]);
					seg += 17;
					if (total > 256 || seg + total > segEnd) return Fail(path, "bad JPEG Huffman table");
					memcpy(t.values, seg, size_t(total));
					seg += total;
					t.Build();
				}
				break;
			case 0xdd: if (length < 4) return Fail(path, "bad JPEG restart interval segment"); restartInterval = int(Be16(seg)); break;
			case 0xee: if (length >= 14 && 0 == memcmp(seg, "Adobe", 5)) adobeTransform = seg[11]; break;
			case 0xc0: case 0xc1: case 0xc2: // SOF0/1 (sequential Huffman), SOF2 (progressive Huffman)
			{
				if (haveFrame) return Fail(path, "more than one JPEG frame");
				progressive = 0xc2 == marker;
				if (length < 8) return Fail(path, "truncated JPEG frame header");
				if (8 != seg[0]) return Fail(path, "only 8-bit JPEG samples are supported");
				height = int(Be16(seg + 1)); width = int(Be16(seg + 3));
				const int n = seg[5];
				if (0 == width || 0 == height || (1 != n && 3 != n) || length < size_t(8 + 3*n)) return Fail(path, "unsupported JPEG frame (grey or 3 components expected)");
				if (width > 16384 || height > 16384) return Fail(path, "JPEG larger than 16384 pixels a side");
				comps.resize(size_t(n));
				for (int i = 0; i < n; ++i)
				{
					Component &c = comps[size_t(i)];
					c.id = seg[6 + i*3]; c.h = seg[7 + i*3] >> 4; c.v = seg[7 + i*3] & 15; c.tq = seg[8 + i*3] & 3;
					if (c.h < 1 || c.h > 2 || c.v < 1 || c.v > 2) return Fail(path, "unsupported JPEG sampling factors");
					if (1 == n) c.h = c.v = 1; // a single component is never interleaved: its sampling factors mean nothing (T.81 A.2.2)
					hMax = c.h > hMax ? c.h : hMax; vMax = c.v > vMax ? c.v : vMax;
				}
				mcusX = (width + 8*hMax - 1)/(8*hMax); mcusY = (height + 8*vMax - 1)/(8*vMax);
				for (Component &c : comps)
				{
					c.blocksW = mcusX*c.h; c.blocksH = mcusY*c.v;
					const int cw = (width*c.h + hMax - 1)/hMax, ch = (height*c.v + vMax - 1)/vMax;
					c.widthInBlocks = (cw + 7)/8; c.heightInBlocks = (ch + 7)/8;
					c.coef.assign(size_t(c.blocksW)*c.blocksH*64, 0);
				}
				haveFrame = true;
				break;
			}
			case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf:
				return Fail(path, "unsupported JPEG process (lossless, hierarchical or arithmetic)");
			case 0xda: // SOS
			{
				if (!haveFrame) return Fail(path, "JPEG scan before the frame header");
				if (length < 3) return Fail(path, "bad JPEG scan header");
				const int n = seg[0];
				if (n < 1 || n > int(comps.size()) || length < size_t(6 + 2*n)) return Fail(path, "bad JPEG scan header");
				std::vector<int> scanComps;
				for (int i = 0; i < n; ++i)
				{
					int found = -1;
					for (size_t c = 0; c < comps.size(); ++c) if (comps[c].id == seg[1 + i*2]) found = int(c);
					if (found < 0) return Fail(path, "JPEG scan names an unknown component");
					comps[size_t(found)].dcTable = (seg[2 + i*2] >> 4) & 3;
					comps[size_t(found)].acTable = seg[2 + i*2] & 3;
					scanComps.push_back(found);
				}
				const int ss = seg[1 + n*2], se = seg[2 + n*2], ah = seg[3 + n*2] >> 4, al = seg[3 + n*2] & 15;
				if (progressive ? (ss > se || se > 63 || (0 == ss && 0 != se) || (ss > 0 && 1 != n)) : false) return Fail(path, "bad JPEG progressive scan");
				for (int ci : scanComps)
				{
					const Component &c = comps[size_t(ci)];
					if ((!progressive || 0 == ss) && (!progressive || 0 == ah) && !dc[c.dcTable].present) return Fail(path, "JPEG scan uses a missing DC table");
					if ((!progressive || ss > 0) && !ac[c.acTable].present) return Fail(path, "JPEG scan uses a missing AC table");
				}
				if (!Scan(pos, scanComps, progressive ? ss : 0, progressive ? se : 63, progressive ? ah : 0, progressive ? al : 0)) return false;
				break;
			}
			default: break; // APPn, COM, ...
			}
		}
		if (!haveFrame) return Fail(path, "JPEG frame header missing");

		// inverse DCT of every block
		for (Component &c : comps)
		{
			const size_t stride = size_t(c.blocksW)*8;
			c.plane.resize(stride*c.blocksH*8);
			for (int by = 0; by < c.blocksH; ++by)
				for (int bx = 0; bx < c.blocksW; ++bx)
					Idct(&c.coef[(size_t(by)*c.blocksW + bx)*64], quant[c.tq], &c.plane[size_t(by)*8*stride + size_t(bx)*8], stride);
		}

		out.width = width; out.height = height;
		if (1 == comps.size())
		{
			out.channels = 1;
			out.px.resize(size_t(width)*height);
			const size_t stride = size_t(comps[0].blocksW)*8;
			for (int y = 0; y < height; ++y) memcpy(&out.px[size_t(y)*width], &comps[0].plane[size_t(y)*stride], size_t(width));
			return true;
		}

		// chroma to full resolution: IJG jdsample.c h2v1/h2v2 "fancy" (triangle) upsampling, or a plain copy at 1x1
		std::vector<uint8_t> full[3];
		for (int i = 0; i < 3; ++i)
		{
			const Component &c = comps[size_t(i)];
			const size_t stride = size_t(c.blocksW)*8;
			full[i].resize(size_t(width)*height);
			const int cw = (width*c.h + hMax - 1)/hMax, ch = (height*c.v + vMax - 1)/vMax;
			if (c.h == hMax && c.v == vMax)
				for (int y = 0; y < height; ++y) memcpy(&full[i][size_t(y)*width], &c.plane[size_t(y)*stride], size_t(width));
			else if (c.h*2 == hMax && c.v == vMax) // h2v1
				for (int y = 0; y < height; ++y)
				{
					const uint8_t *in = &c.plane[size_t(y)*stride];
					uint8_t *o = &full[i][size_t(y)*width];
					for (int x = 0; x < width; ++x)
					{
						const int cx = x >> 1, cur = in[cx];
						if (cw <= 2) { o[x] = uint8_t(cur); continue; } // jdsample.c: fancy only when downsampled_width > 2, else replication
						if (x & 1) o[x] = uint8_t(cx + 1 < cw ? (cur*3 + in[cx + 1] + 2) >> 2 : cur);
						else o[x] = uint8_t(cx > 0 ? (cur*3 + in[cx - 1] + 1) >> 2 : cur);
					}
				}
			else if (c.h*2 == hMax && c.v*2 == vMax) // h2v2
				for (int y = 0; y < height; ++y)
				{
					const int cy = y >> 1;
					const int ny = (y & 1) ? (cy + 1 < ch ? cy + 1 : cy) : (cy > 0 ? cy - 1 : cy);
					const uint8_t *in0 = &c.plane[size_t(cy)*stride], *in1 = &c.plane[size_t(ny)*stride];
					uint8_t *o = &full[i][size_t(y)*width];
					for (int x = 0; x < width; ++x)
					{
						const int cx = x >> 1;
						if (cw <= 2) { o[x] = in0[cx]; continue; }       // plain h2v2 replication (jdsample.c)
						const int cur = in0[cx]*3 + in1[cx];
						if (x & 1) { const int nx = cx + 1 < cw ? in0[cx + 1]*3 + in1[cx + 1] : -1; o[x] = uint8_t(nx >= 0 ? (cur*3 + nx + 7) >> 4 : (cur*4 + 7) >> 4); }
						else { const int lx = cx > 0 ? in0[cx - 1]*3 + in1[cx - 1] : -1; o[x] = uint8_t(lx >= 0 ? (cur*3 + lx + 8) >> 4 : (cur*4 + 8) >> 4); }
					}
				}
			else
				return Fail(path, "unsupported JPEG chroma subsampling");
		}

		// IJG jdcolor.c: YCbCr -> RGB with 16-bit fixed-point tables (Adobe transform 0 = the components already are RGB)
		out.channels = 3;
		out.px.resize(size_t(width)*height*3);
		const bool isRgb = 0 == adobeTransform;
		auto fix = [](double x) { return long(x*65536.0 + 0.5); };
		const long kCrR = fix(1.40200), kCbB = fix(1.77200), kCrG = -fix(0.71414), kCbG = -fix(0.34414), kHalf = 1L << 15;
		auto clamp8 = [](long v) { return uint8_t(v < 0 ? 0 : (v > 255 ? 255 : v)); };
		for (size_t i = 0, n = size_t(width)*height; i < n; ++i)
		{
			const int y = full[0][i];
			if (isRgb) { out.px[i*3] = uint8_t(y); out.px[i*3 + 1] = full[1][i]; out.px[i*3 + 2] = full[2][i]; continue; }
			const long cb = long(full[1][i]) - 128, cr = long(full[2][i]) - 128;
			out.px[i*3]     = clamp8(y + ((kCrR*cr + kHalf) >> 16));
			out.px[i*3 + 1] = clamp8(y + ((kCbG*cb + kCrG*cr + kHalf) >> 16));
			out.px[i*3 + 2] = clamp8(y + ((kCbB*cb + kHalf) >> 16));
		}
		return true;
	}
};

// ---------------------------------------------------------------------------------------------------------------
// file -> the reference's pixel formats
// ---------------------------------------------------------------------------------------------------------------

bool DecodeFile(const std::string &path, Decoded &img)
{
	std::vector<uint8_t> file;
	if (!ReadFile(path, file))
		return Fail(path, "file not found or empty");
	if (file.size() >= 2 && 0xff == file[0] && 0xd8 == file[1])
	{
		JpegDecoder decoder(path, file);
		return decoder.Decode(img);
	}
	return DecodePng(path, file, img);
}

inline uint8_t Luminance(unsigned r, unsigned g, unsigned b) { return uint8_t((r*19595u + g*38470u + b*7471u + 0x8000u) >> 16); } // ITU-R 601 in 16.16

void *ConvertAndAlign(const Decoded &img, bool isGrayscale)
{
	const size_t n = size_t(img.width)*img.height;
	void *pixels = nullptr;
	if (0 != posix_memalign(&pixels, 64, (isGrayscale ? n : n*4) + 64)) // kAlignTo-style alignment, image.cpp:50,57
		return nullptr;
	const uint8_t *s = img.px.data();
	const int ch = img.channels;
	if (isGrayscale)
	{
		uint8_t *d = static_cast<uint8_t *>(pixels);
		for (size_t i = 0; i < n; ++i, s += ch)
			d[i] = ch < 3 ? s[0] : Luminance(s[0], s[1], s[2]);
	}
	else
	{
		uint32_t *d = static_cast<uint32_t *>(pixels);
		for (size_t i = 0; i < n; ++i, s += ch)
		{
			const uint32_t r = s[0], g = ch < 3 ? s[0] : s[1], b = ch < 3 ? s[0] : s[2];
			const uint32_t a = (2 == ch) ? s[1] : (4 == ch ? s[3] : 255u);
			d[i] = (a << 24) | (r << 16) | (g << 8) | b; // IL_BGRA bytes, image.cpp:52-53
		}
	}
	return pixels;
}

void *Load(const std::string &path, bool isGrayscale, int *pWidth, int *pHeight, bool noGC)
{
	Decoded img;
	if (!DecodeFile(path, img))
		return nullptr;
	void *pixels = ConvertAndAlign(img, isGrayscale);
	if (nullptr == pixels) { Fail(path, "out of memory"); return nullptr; }
	if (!noGC) s_gc.push_back(pixels);
	if (pWidth) *pWidth = img.width;
	if (pHeight) *pHeight = img.height;
	return pixels;
}

} // namespace

void CkdHost_SetAssetRoot(const char *directory) { s_assetRoot = directory ? directory : ""; }

// image.cpp:13-30
bool Image_Create() { s_gc.clear(); return true; }

void Image_Destroy()
{
	for (void *p : s_gc) free(p);
	s_gc.clear();
}

uint32_t *Image_Load32(const std::string &path) { return static_cast<uint32_t *>(Load(path, false, nullptr, nullptr, false)); }
uint8_t *Image_Load8(const std::string &path) { return static_cast<uint8_t *>(Load(path, true, nullptr, nullptr, false)); }

// image.cpp:83-110
uint32_t *Image_Load32_CA(const std::string &pathC, const std::string &pathA)
{
	int width = 0, height = 0, widthA = 0, heightA = 0;
	uint32_t *pColor = static_cast<uint32_t *>(Load(pathC, false, &width, &height, false));
	if (nullptr == pColor)
		return nullptr;
	uint32_t *pAlpha = static_cast<uint32_t *>(Load(pathA, false, &widthA, &heightA, true));
	if (nullptr == pAlpha)
		return nullptr;
	const size_t numPixels = size_t(width)*height, numAlpha = size_t(widthA)*heightA;
	for (size_t i = 0; i < numPixels && i < numAlpha; ++i)
		pColor[i] = (pColor[i] & 0xffffff) | (pAlpha[i] & 0xff) << 24;
	free(pAlpha);
	return pColor;
}

namespace ckdhost
{
	// decode `path` (relative to the asset root) for the registry: bpp 4 = BGRA, 1 = L8
	bool DecodeImageFile(const char *path, int bpp, std::vector<uint8_t> &pixels, int &width, int &height)
	{
		Decoded img;
		if (!DecodeFile(path, img))
			return false;
		void *p = ConvertAndAlign(img, 1 == bpp);
		if (nullptr == p) return Fail(path, "out of memory");
		const size_t bytes = size_t(img.width)*img.height*(1 == bpp ? 1 : 4);
		pixels.assign(static_cast<uint8_t *>(p), static_cast<uint8_t *>(p) + bytes);
		free(p);
		width = img.width; height = img.height;
		return true;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// resolution rules ("deterministic 4K resampling rules"; the Python twin is cookiedough_b200/assets.py)
// ---------------------------------------------------------------------------------------------------------------
// The reference's art is made for 1280x720.  For any other output resolution: art of exactly the native output size is
// nearest-resampled to the output size, the two 644x364 blur maps (native FX-map size) to the FX-map size, the 2160x720
// ribbon strip scales with resY/720 (the part-12 strided read stays inside it); textures, sprites and the
// 1280x568 credit logos keep their size.  Nearest means source index = (i*srcSize)/dstSize.  The tunnelscape colour map
// is listed in the reference's .MISSING_LARGE_BLOBS: when the file is absent it is the landscape colour map at twice the
// size, at every resolution (the stand-in both sides of the parity harness use).

namespace {

void NearestResize(std::vector<uint8_t> &pixels, int &width, int &height, int bpp, int newW, int newH)
{
	if (newW == width && newH == height)
		return;
	std::vector<uint8_t> out(size_t(newW)*newH*bpp);
	for (int y = 0; y < newH; ++y)
	{
		const uint8_t *srcRow = &pixels[size_t((int64_t(y)*height)/newH)*width*bpp];
		uint8_t *dstRow = &out[size_t(y)*newW*bpp];
		for (int x = 0; x < newW; ++x)
			memcpy(dstRow + size_t(x)*bpp, srcRow + size_t((int64_t(x)*width)/newW)*bpp, size_t(bpp));
	}
	pixels.swap(out);
	width = newW; height = newH;
}

} // namespace

namespace ckdhost
{
	bool DecodeImageForResolution(const char *path, int bpp, int resX, int resY, std::vector<uint8_t> &pixels, int &width, int &height)
	{
		static const char kMissingMap[] = "assets/scape/tscape-C7W-edit.png", kStandIn[] = "assets/scape/C17W-edit.png";
		if (!DecodeImageFile(path, bpp, pixels, width, height))
		{
			if (0 != strcmp(path, kMissingMap) || !DecodeImageFile(kStandIn, bpp, pixels, width, height))
				return false;
			NearestResize(pixels, width, height, bpp, width*2, height*2);
		}
		constexpr int kNativeX = 1280, kNativeY = 720, kNativeFxX = kNativeX/2 + 4, kNativeFxY = kNativeY/2 + 4;
		if (width == kNativeX && height == kNativeY)
			NearestResize(pixels, width, height, bpp, resX, resY);
		else if (width == kNativeFxX && height == kNativeFxY)
			NearestResize(pixels, width, height, bpp, resX/2 + 4, resY/2 + 4);
		else if (0 == strcmp(path, "assets/demo/ribbons.png"))
			NearestResize(pixels, width, height, bpp, width*resY/kNativeY, height*resY/kNativeY);
		return true;
	}
}

extern "C" {

// ctypes hook: decode a file for an output resolution (rules above) into a malloc'ed buffer (ckdhost_image_free)
void *ckdhost_image_load_for_resolution(const char *path, int grayscale, int resX, int resY, int *width, int *height)
{
	std::vector<uint8_t> pixels;
	int w = 0, h = 0;
	if (!ckdhost::DecodeImageForResolution(path, grayscale ? 1 : 4, resX, resY, pixels, w, h))
		return nullptr;
	void *p = malloc(pixels.size() + 16);
	if (nullptr == p) return nullptr;
	memcpy(p, pixels.data(), pixels.size());
	if (width) *width = w;
	if (height) *height = h;
	return p;
}

// ctypes hook: decode a file; returns a malloc'ed buffer the caller releases with ckdhost_image_free (or null)
void *ckdhost_image_load(const char *path, int grayscale, int *width, int *height) { return Load(path, 0 != grayscale, width, height, true); }
void ckdhost_image_free(void *pixels) { free(pixels); }
void ckdhost_set_asset_root(const char *directory) { CkdHost_SetAssetRoot(directory); }

}
