// cavit-sm100 — input staging for real volumes: stored NIfTI voxels -> the fp32 batch the models take.
//
// The reference's deterministic input pipeline (/root/reference/dataset_ucsf.py:81-89,121-134,157-158) is, per volume,
//   LoadImaged(nibabel)            stored voxels * scl_slope + scl_inter, computed in float64, array index [i][j][k] with
//                                  i the FASTEST axis of the file (Fortran order);
//   EnsureChannelFirstd            a leading channel axis;
//   ResizeWithPadOrCropd(img_size, constant_values=-1)   centre crop, then symmetric constant pad, per axis;
//   .to(torch.float)               C-contiguous fp32 [1][D][H][W] (k fastest).
// On the host that is a float64 temporary plus a strided transposing copy per volume, and the batch crosses PCIe as fp32.
// Here the host only inflates the file and hands over the STORED bytes (int16 for the UCSF volumes: half the H2D bytes);
// one launch per batch scales, crops / pads and transposes every volume of the batch into its slot of the
// [B][M][1][D][H][W] tensor. HBM-bound index work: each stored voxel inside the crop window is read once, coalesced along
// the file's fastest axis, each output voxel is written once, coalesced along the tensor's fastest axis (128 x 32 tiles
// through shared memory; the tile plane is (i, k), or (i, j) for single-slice targets with W == 1).
//
// Arithmetic is bit-exact against the numpy restatement (oracle/staging.py): two separately rounded float64 operations
// (no FMA contraction: nibabel multiplies, then adds), then one rounding to fp32.
#include "common.cuh"
#include "internal.h"

namespace cavit {

struct AxisMap {
  int lo, hi, shift;  // output index o maps to stored index o + shift when lo <= o < hi, to the pad value otherwise
};

// ResizeWithPadOrCrop along one axis: CenterSpatialCrop (start = S/2 - T/2 when S > T), then SpatialPad("symmetric")
// (before = (T - S)/2, the odd voxel goes after).
__device__ __forceinline__ AxisMap axis_map(int S, int T) {
  AxisMap m;
  const int eff = S < T ? S : T;
  const int crop = S > T ? S / 2 - T / 2 : 0;
  m.lo = (T - eff) / 2;
  m.hi = m.lo + eff;
  m.shift = crop - m.lo;
  return m;
}

struct VolDev {  // descriptor with the base pointer resolved
  const uint8_t* raw_base;
  int dims[3];
  float slope, inter;
};

constexpr int STAGE_TI = 128;  // tile extent along i (the file's fastest axis); 32 along the tensor's fastest axis

// Fill the tile with 16 independent loads per thread (four 32-wide i segments x four rows): the kernel is latency-bound on
// 64-byte warp loads otherwise.
template <typename T, bool PLANE_J>
__device__ __forceinline__ void fill_tile(float (*tile)[STAGE_TI + 1], const VolDev& d, const AxisMap mi,
                                          const AxisMap ma, const AxisMap mb, int i0, int a0, int b, float pad) {
  const T* base = reinterpret_cast<const T*>(d.raw_base);
  const bool b_in = b >= mb.lo && b < mb.hi;
  const bool scaled = !(d.slope == 1.0f && d.inter == 0.0f);
  const double slope = static_cast<double>(d.slope), inter = static_cast<double>(d.inter);
  const long long X = d.dims[0], XY = X * d.dims[1];
  T raw[4][4];
  bool in[4][4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int a = a0 + threadIdx.y + 8 * rr;
    const bool a_in = b_in && a >= ma.lo && a < ma.hi;
    const long long sj = (PLANE_J ? a + ma.shift : b + mb.shift), sk = (PLANE_J ? b + mb.shift : a + ma.shift);
    const long long row = X * sj + XY * sk + mi.shift;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + 32 * q + threadIdx.x;
      in[rr][q] = a_in && i >= mi.lo && i < mi.hi;
      raw[rr][q] = in[rr][q] ? __ldg(base + row + i) : T(0);
    }
  }
#pragma unroll
  for (int rr = 0; rr < 4; ++rr)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v = pad;
      if (in[rr][q]) {
        if (scaled) {
          double x = static_cast<double>(raw[rr][q]);
          if (slope != 1.0) x = __dmul_rn(x, slope);
          if (inter != 0.0) x = __dadd_rn(x, inter);
          v = static_cast<float>(x);
        } else {
          v = static_cast<float>(raw[rr][q]);  // == (float)(double)raw: the widening is exact, one rounding either way
        }
      }
      tile[threadIdx.y + 8 * rr][32 * q + threadIdx.x] = v;
    }
}

// grid = (tiles over (i, a), extent of b, volumes), block = (32, 8); (a, b) = (k, j), or (j, k) when PLANE_J.
template <bool PLANE_J>
__global__ void __launch_bounds__(256)
stage_volumes_kernel(const uint8_t* __restrict__ raw, const cavit_volume_desc* __restrict__ desc, float* __restrict__ out,
                     int D, int H, int W, float pad) {
  __shared__ float tile[32][STAGE_TI + 1];
  const cavit_volume_desc dd = desc[blockIdx.z];
  const VolDev d = {raw + dd.byte_offset, {dd.dims[0], dd.dims[1], dd.dims[2]}, dd.slope, dd.inter};
  const int A = PLANE_J ? H : W;
  const int tilesA = (A + 31) / 32;
  const int i0 = (blockIdx.x / tilesA) * STAGE_TI, a0 = (blockIdx.x % tilesA) * 32;
  const int b = blockIdx.y;
  const AxisMap mi = axis_map(d.dims[0], D);
  const AxisMap mj = axis_map(d.dims[1], H);
  const AxisMap mk = axis_map(d.dims[2], W);
  const AxisMap ma = PLANE_J ? mj : mk, mb = PLANE_J ? mk : mj;
  switch (dd.dtype) {  // block-uniform
    case CAVIT_VOX_U8: fill_tile<uint8_t, PLANE_J>(tile, d, mi, ma, mb, i0, a0, b, pad); break;
    case CAVIT_VOX_I16: fill_tile<int16_t, PLANE_J>(tile, d, mi, ma, mb, i0, a0, b, pad); break;
    case CAVIT_VOX_I32: fill_tile<int32_t, PLANE_J>(tile, d, mi, ma, mb, i0, a0, b, pad); break;
    case CAVIT_VOX_F32: fill_tile<float, PLANE_J>(tile, d, mi, ma, mb, i0, a0, b, pad); break;
    case CAVIT_VOX_F64: fill_tile<double, PLANE_J>(tile, d, mi, ma, mb, i0, a0, b, pad); break;
    case CAVIT_VOX_I8: fill_tile<int8_t, PLANE_J>(tile, d, mi, ma, mb, i0, a0, b, pad); break;
    case CAVIT_VOX_U16: fill_tile<uint16_t, PLANE_J>(tile, d, mi, ma, mb, i0, a0, b, pad); break;
    default: fill_tile<uint32_t, PLANE_J>(tile, d, mi, ma, mb, i0, a0, b, pad); break;
  }
  __syncthreads();
  float* o = out + static_cast<long long>(blockIdx.z) * D * H * W;
  const int a = a0 + threadIdx.x;
  const long long j = PLANE_J ? a : b, k = PLANE_J ? b : a;
#pragma unroll 4
  for (int r = threadIdx.y; r < STAGE_TI; r += 8) {
    const int io = i0 + r;
    if (io < D && a < A) o[(static_cast<long long>(io) * H + j) * W + k] = tile[threadIdx.x][r];
  }
}

}  // namespace cavit

using namespace cavit;

extern "C" int cavit_stage_volumes(const void* raw, const cavit_volume_desc* desc, float* out, int32_t volumes, int32_t D,
                                   int32_t H, int32_t W, float pad_value, void* stream) {
  if (!raw || !desc || !out) return fail(CAVIT_E_BADARG, "cavit_stage_volumes: null pointer");
  if (volumes < 1 || D < 1 || H < 1 || W < 1) return fail(CAVIT_E_BADARG, "cavit_stage_volumes: empty batch or target");
  if (volumes > 65535 || H > 65535 || W > 65535)
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_stage_volumes: more than 65535 volumes, rows or columns");
  const bool plane_j = (W == 1);
  const int A = plane_j ? H : W, Bx = plane_j ? W : H;
  const long long tiles = static_cast<long long>((D + STAGE_TI - 1) / STAGE_TI) * ((A + 31) / 32);
  if (tiles > 2147483647LL) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_stage_volumes: target too large");
  const dim3 grid(static_cast<unsigned>(tiles), Bx, volumes), block(32, 8);
  if (plane_j)
    stage_volumes_kernel<true><<<grid, block, 0, as_stream(stream)>>>(static_cast<const uint8_t*>(raw), desc, out, D, H, W,
                                                                      pad_value);
  else
    stage_volumes_kernel<false><<<grid, block, 0, as_stream(stream)>>>(static_cast<const uint8_t*>(raw), desc, out, D, H, W,
                                                                       pad_value);
  count_launch();
  return check_launch("cavit_stage_volumes");
}
