// Row N2 of SURVEY.md section 8f: hand-over of the lifted volume to FastIndoorImVoxelNeck (necks/imvoxelnet.py:8-67).
//
// The lift writes alpha * mean as channel-major rows [C][X*Y*Z] fp32 (the reference's layout, nerfdet.py:259-266); the neck's
// first Conv3d runs fastest in cuDNN on channels-last-3D bf16 input.  Instead of torch's stack + to(bf16) +
// contiguous(channels_last_3d) (three passes over 26 MB), one tiled transpose reads the rows once and writes
// [X*Y*Z][C] in the requested dtype; the same launch turns the int64 view counts into the float `valids` the head
// up-samples (imvoxel_head_v2.py:93, nerfdet.py:287).
#include "nd_common.cuh"

namespace nd {

template <typename T> __device__ __forceinline__ T neck_cast(float v);
template <> __device__ __forceinline__ float neck_cast<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 neck_cast<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// block (32, 8): tile of 32 voxels x 32 channels through shared memory
template <typename T>
__global__ void k_volume_to_neck(const float *__restrict__ vol, const int64_t *__restrict__ count, int channels, int64_t n_vox,
                                 T *__restrict__ out, float *__restrict__ valid) {
    __shared__ float tile[32][33];
    const int64_t n0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j;
        const int64_t n = n0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < channels && n < n_vox) ? vol[(int64_t)c * n_vox + n] : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int64_t n = n0 + j;
        const int c = c0 + threadIdx.x;
        if (n < n_vox && c < channels) out[n * channels + c] = neck_cast<T>(tile[threadIdx.x][j]);
    }
    if (valid != nullptr && blockIdx.y == 0 && threadIdx.y == 0 && n0 + threadIdx.x < n_vox)
        valid[n0 + threadIdx.x] = (float)count[n0 + threadIdx.x];
}

}  // namespace nd

using namespace nd;

extern "C" {

int nd_volume_to_neck(const float *volume, const int64_t *count, int channels, int64_t n_voxels, int out_dtype, void *out,
                      float *valid, void *stream) {
    ND_REQUIRE(volume && out && (count || !valid), ND_ERR_BAD_ARG, "nd_volume_to_neck: null pointer");
    ND_REQUIRE(channels > 0 && n_voxels >= 0, ND_ERR_BAD_SHAPE, "nd_volume_to_neck: bad shape");
    ND_REQUIRE(out_dtype == ND_F32 || out_dtype == ND_BF16, ND_ERR_BAD_ARG, "nd_volume_to_neck: dtype");
    if (n_voxels == 0) return ND_OK;
    const dim3 grid((unsigned)ceil_div(n_voxels, 32), (unsigned)ceil_div(channels, 32)), block(32, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == ND_F32)
        k_volume_to_neck<float><<<grid, block, 0, st>>>(volume, count, channels, n_voxels, (float *)out, valid);
    else
        k_volume_to_neck<__nv_bfloat16><<<grid, block, 0, st>>>(volume, count, channels, n_voxels, (__nv_bfloat16 *)out, valid);
    ND_CUDA_LAUNCH_CHECK("k_volume_to_neck");
    return ND_OK;
}

}  // extern "C"
