// Tile post-processing kernels: normalise + tone-map + 8-bit pack of every wall's lightmap on the
// device, optionally straight into complete PNG files.  Not part of the trace path (and of its source hash).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fmgi {

// ---- tile post-processing ---------------------------------------------------------------
//
// Device version of what the caller does to the lightmap before it becomes tiles/tile_N.png:
// main.c:68-79 (texel *= 0.35 * tiles / (area * samplesPerArea)) and saveAs_core (rectangle.c:293-336:
// tone-map 1 - exp(-2 L) at constant chroma, x255, clamp, floor tint).  One thread per base-level
// texel; every implicit promotion of the reference is kept (double luminance sum and exp, float
// ratio, uint8 x double floor tint) with explicitly rounded operations so that nothing is contracted.
struct TileWall {
    int32_t base;        // atlas index of the wall's base level
    int32_t first;       // index of its first texel in the packed RGB output
    float scale;         // (float)(0.35 * tilesPerSample), main.c:73-77
    int32_t is_floor;    // rectangle.c:317
};

__device__ __forceinline__ unsigned char tile_clamp(float d)          // rectangle.c:287-292
{
    if (d < 0.0f) d = 0.0f;
    if (d > 255.0f) d = 255.0f;
    return (unsigned char)__float2uint_rz(d);                          // NaN (black texel, 0/0) -> 0
}

// PNG layout of one tile (8-bit RGB, filter 0, zlib stream of STORED blocks - no compression, so every byte of the
// file has a position known in advance): signature 8 | IHDR chunk 25 | IDAT length + type 8 | zlib header 2 |
// stored blocks of at most 65535 raw bytes, 5 bytes of header each | Adler-32 4 | IDAT CRC 4 | IEND chunk 12.
// Raw data = height rows of (1 filter byte + 3 * width pixel bytes).
struct PngWall {
    long long file_off;  // offset of the tile's file in the output buffer
    int32_t width, height;
};
constexpr int kPngDataStart = 8 + 25 + 8 + 2;      // file offset of the first stored-block header
constexpr unsigned kPngBlock = 65535u;
__host__ __device__ __forceinline__ unsigned long long png_raw_bytes(int w, int h) { return (unsigned long long)h * (3ull * w + 1ull); }
__host__ __device__ __forceinline__ unsigned long long png_file_bytes(int w, int h)
{
    const unsigned long long raw = png_raw_bytes(w, h);
    const unsigned long long blocks = raw ? (raw + kPngBlock - 1) / kPngBlock : 1;
    return 8 + 25 + 12 + (2 + 5 * blocks + raw + 4) + 12;
}
// file offset of raw byte r
__device__ __forceinline__ unsigned long long png_pos(unsigned long long r) { return kPngDataStart + 5ull * (r / kPngBlock + 1ull) + r; }

// kPng = false: packed RGB in wall order (rgb[3 * pixel]); kPng = true: the pixel bytes go to their place inside the
// wall's PNG file (k_png_finish writes everything around them).
template <bool kPng>
__global__ void k_tonemap(const float4 *__restrict__ atlas, const TileWall *__restrict__ walls, int num_walls,
                          long long num_pixels, int tint_extra, unsigned char *__restrict__ rgb,
                          const PngWall *__restrict__ png)
{
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < num_pixels;
         g += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = num_walls - 1;                                // wall whose pixel range holds g
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((long long)walls[mid].first <= g) lo = mid; else hi = mid - 1;
        }
        const TileWall w = walls[lo];
        const float4 t = atlas[w.base + (int)(g - w.first)];
        float r = __fmul_rn(t.x, w.scale), gg = __fmul_rn(t.y, w.scale), b = __fmul_rn(t.z, w.scale);
        const float lum = (float)__dadd_rn(__dadd_rn(__dmul_rn(0.2126, (double)r), __dmul_rn(0.7152, (double)gg)),
                                           __dmul_rn(0.0722, (double)b));                  // rectangle.c:277
        const float perceptive = (float)__dsub_rn(1.0, exp((double)__fmul_rn(-2.0f, lum)));   // rectangle.c:269
        const float q = __fdiv_rn(perceptive, lum);
        r = __fmul_rn(r, q); gg = __fmul_rn(gg, q); b = __fmul_rn(b, q);
        unsigned char d0 = tile_clamp(__fmul_rn(r, 255.0f)), d1 = tile_clamp(__fmul_rn(gg, 255.0f)),
                      d2 = tile_clamp(__fmul_rn(b, 255.0f));
        if (w.is_floor) {                                              // rectangle.c:317-334
            d1 = (unsigned char)__double2uint_rz(__dmul_rn((double)d1, 0.95));
            d2 = (unsigned char)__double2uint_rz(__dmul_rn((double)d2, 0.9));
            if (tint_extra) {
                d1 = (unsigned char)__float2uint_rz(__fmul_rn((float)d1, 0.95f));
                d2 = (unsigned char)__float2uint_rz(__fmul_rn((float)d2, 0.9f));
            }
        }
        if (!kPng) {
            rgb[3 * g] = d0; rgb[3 * g + 1] = d1; rgb[3 * g + 2] = d2;
        } else {
            const PngWall pw = png[lo];
            const int j = (int)(g - w.first), x = j % pw.width, y = j / pw.width;
            const unsigned long long raw = (unsigned long long)y * (3ull * pw.width + 1ull) + 1ull + 3ull * x;
            unsigned char *f = rgb + pw.file_off;
            f[png_pos(raw)] = d0; f[png_pos(raw + 1)] = d1; f[png_pos(raw + 2)] = d2;
        }
    }
}

// Everything of a tile's PNG file but the pixel bytes: one thread per wall writes the signature, IHDR, the IDAT
// framing (zlib header, stored-block headers, the rows' filter bytes), then runs Adler-32 over the raw data and
// CRC-32 over the chunks (table in shared memory) and appends IEND.  write_png_file (png_helper.c:255) produces the
// same pixels with libpng's deflate; a decoder sees identical images.
__device__ __forceinline__ void png_put32(unsigned char *p, unsigned v)
{
    p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v;
}

__global__ void __launch_bounds__(256) k_png_finish(const PngWall *__restrict__ png, int num_walls, unsigned char *__restrict__ out)
{
    __shared__ unsigned crc_table[256];
    {
        unsigned c = threadIdx.x;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crc_table[threadIdx.x] = c;
    }
    __syncthreads();
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= num_walls) return;
    const PngWall pw = png[wi];
    unsigned char *f = out + pw.file_off;
    auto crc_bytes = [&](unsigned crc, const unsigned char *p, unsigned long long n) {
        for (unsigned long long i = 0; i < n; i++) crc = crc_table[(crc ^ p[i]) & 255u] ^ (crc >> 8);
        return crc;
    };
    // signature + IHDR
    const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    for (int i = 0; i < 8; i++) f[i] = sig[i];
    png_put32(f + 8, 13u);
    f[12] = 'I'; f[13] = 'H'; f[14] = 'D'; f[15] = 'R';
    png_put32(f + 16, (unsigned)pw.width); png_put32(f + 20, (unsigned)pw.height);
    f[24] = 8; f[25] = 2; f[26] = 0; f[27] = 0; f[28] = 0;             // 8 bits, colour type RGB, deflate, filter 0, no interlace
    png_put32(f + 29, ~crc_bytes(0xffffffffu, f + 12, 17));
    // IDAT framing
    const unsigned long long raw = png_raw_bytes(pw.width, pw.height);
    const unsigned long long blocks = raw ? (raw + kPngBlock - 1) / kPngBlock : 1;
    const unsigned long long zlen = 2 + 5 * blocks + raw + 4;
    png_put32(f + 33, (unsigned)zlen);
    f[37] = 'I'; f[38] = 'D'; f[39] = 'A'; f[40] = 'T';
    f[41] = 0x78; f[42] = 0x01;                                        // zlib: deflate, 32K window, no preset, fastest
    for (unsigned long long b = 0; b < blocks; b++) {
        const unsigned long long left = raw - b * kPngBlock;
        const unsigned len = (unsigned)(left < kPngBlock ? left : kPngBlock);
        unsigned char *h = f + kPngDataStart + b * (kPngBlock + 5ull);
        h[0] = b + 1 == blocks ? 1 : 0;                                // BFINAL, BTYPE = 00 (stored)
        h[1] = (unsigned char)len; h[2] = (unsigned char)(len >> 8);
        h[3] = (unsigned char)~len; h[4] = (unsigned char)(~len >> 8);
    }
    const unsigned long long stride = 3ull * pw.width + 1ull;
    for (int y = 0; y < pw.height; y++) f[png_pos((unsigned long long)y * stride)] = 0;      // filter type none
    // Adler-32 over the raw data (as laid out in the file: block by block), then the CRC of the whole chunk
    unsigned s1 = 1u, s2 = 0u;
    for (unsigned long long b = 0; b < blocks; b++) {
        const unsigned long long left = raw - b * kPngBlock;
        const unsigned len = (unsigned)(left < kPngBlock ? left : kPngBlock);
        const unsigned char *d = f + kPngDataStart + b * (kPngBlock + 5ull) + 5;
        for (unsigned i = 0; i < len;) {
            const unsigned run = len - i < 5552u ? len - i : 5552u;   // largest run without overflowing 32 bits
            for (unsigned k = 0; k < run; k++) { s1 += d[i + k]; s2 += s1; }
            s1 %= 65521u; s2 %= 65521u;
            i += run;
        }
    }
    unsigned char *tail = f + kPngDataStart + 5 * blocks + raw;
    png_put32(tail, (s2 << 16) | s1);
    png_put32(tail + 4, ~crc_bytes(0xffffffffu, f + 37, 4 + zlen));
    // IEND
    png_put32(tail + 8, 0u);
    tail[12] = 'I'; tail[13] = 'E'; tail[14] = 'N'; tail[15] = 'D';
    png_put32(tail + 16, 0xAE426082u);
}

}  // namespace fmgi
