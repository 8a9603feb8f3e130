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
};

template <typename TS>
__device__ __forceinline__ float ld(const TS* p, long i);
template <>
__device__ __forceinline__ float ld<float>(const float* p, long i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float ld<__half>(const __half* p, long i) { return __half2float(__ldg(p + i)); }

template <typename TS, typename TD>
__device__ __forceinline__ void unpack_tensor(const TS* __restrict__ src, TD* __restrict__ dst, int C, int h, int w, int code,
                                              int chunk, int nchunks) {
  const bool tr = code & 1, fy = code & 2, fx = code & 4;
  const int Ho = tr ? w : h, Wo = tr ? h : w;
  const long plane = (long)h * w;
  if ((Wo & 3) == 0) {
    const int wq = Wo >> 2;
    const long total = (long)C * Ho * wq;
    for (long i = (long)chunk * blockDim.x + threadIdx.x; i < total; i += (long)nchunks * blockDim.x) {
      const int xq = (int)(i % wq);
      const long r = i / wq;
      const int y = (int)(r % Ho);
      const long c = r / Ho;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = 4 * xq + j;
        int sy = tr ? x : y, sx = tr ? y : x;
        if (fy) sy = h - 1 - sy;
        if (fx) sx = w - 1 - sx;
        v[j] = ld<TS>(src, c * plane + (long)sy * w + sx);
      }
      const long o = (c * Ho + y) * (long)Wo + 4 * xq;
      if constexpr (sizeof(TD) == 4) {
        *reinterpret_cast<float4*>(dst + o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(dst + o) = u;
      }
    }
  } else {
    const long total = (long)C * Ho * Wo;
    for (long i = (long)chunk * blockDim.x + threadIdx.x; i < total; i += (long)nchunks * blockDim.x) {
      const int x = (int)(i % Wo);
      const long r = i / Wo;
      const int y = (int)(r % Ho);
      const long c = r / Ho;
      int sy = tr ? x : y, sx = tr ? y : x;
      if (fy) sy = h - 1 - sy;
      if (fx) sx = w - 1 - sx;
      const float v = ld<TS>(src, c * plane + (long)sy * w + sx);
      if constexpr (sizeof(TD) == 4) dst[i] = v;
      else dst[i] = __float2bfloat16_rn(v);
    }
  }
}

// grid = (chunks, segments, samples)
__global__ void __launch_bounds__(256) k_cache_unpack(const unsigned char* __restrict__ records, size_t record_bytes,
                                                      const __grid_constant__ Segs segs, const int* __restrict__ codes) {
  const ffsr_cache_segment& sg = segs.s[blockIdx.y];
  const int b = blockIdx.z;
  const int code = codes ? codes[b] : 0;
  const unsigned char* src = records + (size_t)b * record_bytes + sg.src_offset;
  const long n = (long)sg.C * sg.h * sg.w;
  // small tensors (lr) need few CTAs: only the first `nchunks` of the grid's x-dimension work on this segment
  const int chunk = blockIdx.x;
  const int nchunks = (int)min((long)gridDim.x, (n + 1023) / 1024);
  if (chunk >= nchunks) return;
  if (sg.src_dtype == FFSR_DT_F16) {
    if (sg.dst_dtype == FFSR_DT_BF16)
      unpack_tensor<__half, __nv_bfloat16>((const __half*)src, (__nv_bfloat16*)sg.dst + (size_t)b * n, sg.C, sg.h, sg.w, code, chunk, nchunks);
    else
      unpack_tensor<__half, float>((const __half*)src, (float*)sg.dst + (size_t)b * n, sg.C, sg.h, sg.w, code, chunk, nchunks);
  } else {
    if (sg.dst_dtype == FFSR_DT_BF16)
      unpack_tensor<float, __nv_bfloat16>((const float*)src, (__nv_bfloat16*)sg.dst + (size_t)b * n, sg.C, sg.h, sg.w, code, chunk, nchunks);
    else
      unpack_tensor<float, float>((const float*)src, (float*)sg.dst + (size_t)b * n, sg.C, sg.h, sg.w, code, chunk, nchunks);
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
  long biggest = 0;
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
    // a quarter turn of a non-square tensor changes [h][w] to [w][h]: per-sample planes stay C*h*w, so dense batches work
    sv.s[i] = s;
    biggest = max(biggest, (long)s.C * s.h * s.w);
  }
  // enough CTAs to fill the machine a few times over, no more than one per 1024 elements of the largest tensor
  const long want = (long)(sm_count > 0 ? sm_count : 148) * 8;
  long chunks = (biggest + 1023) / 1024;
  const long per = (want + (long)nseg * B - 1) / ((long)nseg * B);
  if (chunks > per) chunks = per;
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (unsigned)nseg, (unsigned)B);
  k_cache_unpack<<<grid, 256, 0, stream>>>((const unsigned char*)records, record_bytes, sv, tf_codes);
  return ffsr_check_launch("k_cache_unpack");
}
