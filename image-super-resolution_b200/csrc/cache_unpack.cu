// Flat cache-shard records -> the batch tensors the fusion module consumes.
//
// The reference's loader (src/data/cached_dataset.py:135-282) unpickles three files per sample, up-casts the fp16
// MambaIR part, applies hflip / vflip / rot90 on the CPU (:236-282) and lets the DataLoader collate.  Here a batch
// arrives as B raw records in ONE host->device copy; this kernel is the whole "collate": for every tensor of every
// sample it up-casts (fp16 -> fp32, or -> bf16), applies that sample's dihedral transform and writes the dense
// [B][C][Ho][Wo] batch tensor -- one pass over the bytes, one launch per batch.
//
// A composition of flips and quarter turns is one of the 8 dihedral maps: out[y][x] = in[sy][sx] with
// (sy, sx) = (code & 1) ? (x, y) : (y, x), then sy = h-1-sy if (code & 2), sx = w-1-sx if (code & 4).
#include "common.cuh"
#include "../../include/ffsr_b200.h"
#include <cuda_fp16.h>

namespace {

constexpr int MAX_SEGS = 16;
struct Segs {
  ffsr_cache_segment s[MAX_SEGS];
  int tile_start[MAX_SEGS + 1];      // prefix sum of 64x64 tiles per tensor of one sample
};

template <typename TS>
__device__ __forceinline__ float ld(const TS* p, long i);
template <>
__device__ __forceinline__ float ld<float>(const float* p, long i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float ld<__half>(const __half* p, long i) { return __half2float(__ldg(p + i)); }

constexpr int UT = 64;           // tile edge: one warp row of a tile is 128 B (fp16) / 256 B (fp32) of source
constexpr int UP = UT + 1;       // shared-memory row pitch in words: transposed reads are conflict-free

template <typename TD>
__device__ __forceinline__ void st(TD* p, int i, float v);
template <>
__device__ __forceinline__ void st<float>(float* p, int i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void st<__nv_bfloat16>(__nv_bfloat16* p, int i, float v) { p[i] = __float2bfloat16_rn(v); }

// One 64x64 tile of one channel plane.  The source tile is read row by row (coalesced), parked in shared memory and
// written out along the rows of the OUTPUT, whatever the dihedral map: a quarter turn becomes a transposed read of
// shared memory instead of a strided read of HBM.  The map from output-tile coordinates (r, cc) to the shared-memory
// word is affine: idx0 + r * step_r + cc * step_c.
template <typename TS, typename TD>
__device__ __forceinline__ void unpack_tile(const TS* __restrict__ sp, TD* __restrict__ dp, int h, int w, int code, int tyi, int txi,
                                            float* __restrict__ tile) {
  const bool tr = code & 1, fy = code & 2, fx = code & 4;
  const int Wo = tr ? h : w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int y0 = tyi * UT, x0 = txi * UT;
  const int y1 = min(y0 + UT, h), x1 = min(x0 + UT, w);         // source rectangle [y0,y1) x [x0,x1)
  const int th = y1 - y0, tw = x1 - x0;
  const bool c0 = lane < tw, c1 = lane + 32 < tw;
  const TS* row = sp + y0 * w + x0 + lane;
#pragma unroll 4
  for (int r = warp; r < th; r += 8) {
    const TS* q = row + r * w;
    tile[r * UP + lane] = c0 ? ld<TS>(q, 0) : 0.f;
    tile[r * UP + lane + 32] = c1 ? ld<TS>(q, 32) : 0.f;
  }
  __syncthreads();
  // image of the source rectangle in the output: rows [oy0, oy0+nr), columns [ox0, ox0+nc)
  const int u0 = fy ? h - y1 : y0, v0 = fx ? w - x1 : x0;        // first row / column after the reversals
  const int oy0 = tr ? v0 : u0, ox0 = tr ? u0 : v0;
  const int nr = tr ? tw : th, nc = tr ? th : tw;
  // (r, cc) -> (ly, lx): !tr: ly = fy ? th-1-r : r,  lx = fx ? tw-1-cc : cc;   tr: ly = fy ? th-1-cc : cc,  lx = fx ? tw-1-r : r
  const int sy_ = fy ? -UP : UP, sx_ = fx ? -1 : 1;
  const int idx0 = (fy ? (th - 1) * UP : 0) + (fx ? tw - 1 : 0);
  const int step_r = tr ? sx_ : sy_, step_c = tr ? sy_ : sx_;
  const bool d0 = lane < nc, d1 = lane + 32 < nc;
  const float* t0 = tile + idx0 + lane * step_c;
  TD* orow = dp + oy0 * Wo + ox0 + lane;
#pragma unroll 4
  for (int r = warp; r < nr; r += 8) {
    const float* tp = t0 + r * step_r;
    if (d0) st<TD>(orow, r * Wo, tp[0]);
    if (d1) st<TD>(orow, r * Wo + 32, tp[32 * step_c]);
  }
  __syncthreads();
}

// Work items are (sample, tile) pairs of ALL tensors in one flattened list; every CTA takes one contiguous run of it
// (decoded once, then advanced with carries: no per-tile divisions), so the CTAs get equal work whatever the tensor sizes.
__global__ void __launch_bounds__(256) k_cache_unpack(const unsigned char* __restrict__ records, size_t record_bytes,
                                                      const __grid_constant__ Segs segs, int nseg, int B,
                                                      const int* __restrict__ codes) {
  __shared__ float tile[UT * UP];
  const int T = segs.tile_start[nseg];
  const int items = B * T;
  const int per = (items + (int)gridDim.x - 1) / (int)gridDim.x;
  int item = (int)blockIdx.x * per;
  const int end = min(item + per, items);
  if (item >= end) return;
  int b = item / T;
  int t = item - b * T;
  int si = 0;
  while (si + 1 < nseg && t >= segs.tile_start[si + 1]) ++si;
  t -= segs.tile_start[si];
  int h = segs.s[si].h, w = segs.s[si].w, Cc = segs.s[si].C;
  int tyn = (h + UT - 1) / UT, txn = (w + UT - 1) / UT;
  int txi = t % txn, tyi = (t / txn) % tyn, c = t / (txn * tyn);
  int code = codes ? codes[b] : 0;
  for (; item < end; ++item) {
    const ffsr_cache_segment& sg = segs.s[si];
    const int plane = h * w;
    const unsigned char* src = records + (size_t)b * record_bytes + sg.src_offset;
    const size_t o = ((size_t)b * Cc + c) * plane;
    if (sg.src_dtype == FFSR_DT_F16) {
      const __half* sp = (const __half*)src + (size_t)c * plane;
      if (sg.dst_dtype == FFSR_DT_BF16) unpack_tile<__half, __nv_bfloat16>(sp, (__nv_bfloat16*)sg.dst + o, h, w, code, tyi, txi, tile);
      else unpack_tile<__half, float>(sp, (float*)sg.dst + o, h, w, code, tyi, txi, tile);
    } else {
      const float* sp = (const float*)src + (size_t)c * plane;
      if (sg.dst_dtype == FFSR_DT_BF16) unpack_tile<float, __nv_bfloat16>(sp, (__nv_bfloat16*)sg.dst + o, h, w, code, tyi, txi, tile);
      else unpack_tile<float, float>(sp, (float*)sg.dst + o, h, w, code, tyi, txi, tile);
    }
    if (++txi == txn) {
      txi = 0;
      if (++tyi == tyn) {
        tyi = 0;
        if (++c == Cc) {
          c = 0;
          if (++si == nseg) {
            si = 0;
            ++b;
            if (b < B) code = codes ? codes[b] : 0;
          }
          h = segs.s[si].h, w = segs.s[si].w, Cc = segs.s[si].C;
          tyn = (h + UT - 1) / UT, txn = (w + UT - 1) / UT;
        }
      }
    }
  }
}

}  // namespace

extern "C" int ffsr_cache_segment_size(void) { return (int)sizeof(ffsr_cache_segment); }

extern "C" int ffsr_cache_unpack(const void* records, size_t record_bytes, int B, const ffsr_cache_segment* segs, int nseg,
                                 const int* tf_codes, int sm_count, cudaStream_t stream) {
  FFSR_REQUIRE(records && segs && B > 0 && nseg > 0 && nseg <= MAX_SEGS, FFSR_ERR_ARG,
               "cache_unpack: B=%d nseg=%d (1..%d segments)", B, nseg, MAX_SEGS);
  FFSR_REQUIRE((reinterpret_cast<uintptr_t>(records) & 15) == 0 && (record_bytes & 15) == 0, FFSR_ERR_ALIGN,
               "cache_unpack: records / record_bytes must be 16-byte aligned");
  Segs sv;
  int tiles = 0;
  for (int i = 0; i < nseg; ++i) {
    const ffsr_cache_segment& s = segs[i];
    FFSR_REQUIRE(s.dst && s.C > 0 && s.h > 0 && s.w > 0, FFSR_ERR_ARG, "cache_unpack: segment %d has a null dst or empty shape", i);
    FFSR_REQUIRE(s.src_dtype == FFSR_DT_F32 || s.src_dtype == FFSR_DT_F16, FFSR_ERR_ARG, "cache_unpack: segment %d src_dtype %d", i, s.src_dtype);
    FFSR_REQUIRE(s.dst_dtype == FFSR_DT_F32 || s.dst_dtype == FFSR_DT_BF16, FFSR_ERR_ARG, "cache_unpack: segment %d dst_dtype %d", i, s.dst_dtype);
    const size_t esz = s.src_dtype == FFSR_DT_F16 ? 2 : 4;
    FFSR_REQUIRE((s.src_offset & 15) == 0 && s.src_offset + (size_t)s.C * s.h * s.w * esz <= record_bytes, FFSR_ERR_ARG,
                 "cache_unpack: segment %d [%llu, +%zu) leaves the %zu-byte record or is not 16-byte aligned", i,
                 (unsigned long long)s.src_offset, (size_t)s.C * s.h * s.w * esz, record_bytes);
    FFSR_REQUIRE((reinterpret_cast<uintptr_t>(s.dst) & 15) == 0, FFSR_ERR_ALIGN, "cache_unpack: segment %d dst not 16-byte aligned", i);
    FFSR_REQUIRE((long)s.h * s.w < (1L << 30) && (long)s.C * ((s.h + UT - 1) / UT) * ((s.w + UT - 1) / UT) < (1L << 24), FFSR_ERR_ARG,
                 "cache_unpack: segment %d is too large (%d x %d x %d)", i, s.C, s.h, s.w);
    // a quarter turn of a non-square tensor changes [h][w] to [w][h]: per-sample planes stay C*h*w, so dense batches work
    sv.s[i] = s;
    sv.tile_start[i] = tiles;
    tiles += s.C * ((s.h + UT - 1) / UT) * ((s.w + UT - 1) / UT);
  }
  sv.tile_start[nseg] = tiles;
  FFSR_REQUIRE((long)B * tiles < (1L << 30), FFSR_ERR_ARG, "cache_unpack: %d samples x %d tiles is too many for one launch", B, tiles);
  // persistent-style grid: a few CTAs per SM (16.6 KB of shared memory each), never more than there are tiles
  long grid = (long)(sm_count > 0 ? sm_count : 148) * 8;
  if (grid > (long)B * tiles) grid = (long)B * tiles;
  k_cache_unpack<<<(unsigned)grid, 256, 0, stream>>>((const unsigned char*)records, record_bytes, sv, nseg, B, tf_codes);
  return ffsr_check_launch("k_cache_unpack");
}
