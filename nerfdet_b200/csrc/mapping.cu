// B7 of SURVEY.md section 8a: the per-pixel Linear(C -> Cm) of reference nerfdet.py:190-197,
//   mapped[v][pixel][j] = bias[j] + sum_c weight[j][c] * features[v][c][pixel],
// as ONE kernel that reads the NCHW planes in place (the reference's 242 MB permute().contiguous() copy is not
// made) and writes the channels-last layout [nv][h][w][Cm] the gather kernels (live.cu, render.cu) consume.
//
// fp32 FFMA with fp32 accumulation, channels ascending (the 1e-4 path; TF32 stays off like the reference's GEMM).
// A skinny GEMM: M = nv*h*w pixels (236 000), N = Cm = 32, K = C = 256 -> 3.9 GFLOP against 242 MB of input:
// FFMA-bound (54 us at the 72 TFLOP/s fp32 peak), not HBM-bound (37 us).
//
//   CTA      256 pixels of one view x all 32 outputs, 128 threads; thread (lane, warp) owns pixels
//            {4 lane .. 4 lane + 3} and {128 + 4 lane .. 128 + 4 lane + 3} (two conflict-free LDS.128 per channel)
//            x outputs {8 warp .. 8 warp + 7} (two warp-uniform broadcast LDS.128 per channel): 64 accumulators
//   planes   [8 channels][256 pixels] stages through a 4-deep cp.async ring, each row a coalesced 1 KB segment;
//            persistent CTAs (3 per SM, 64 KB of shared memory each): the ring runs across tile boundaries
//   weights  transposed once per CTA into shared memory, Wt[c][j]
#include "nd_common.cuh"

namespace nd {

constexpr int kMapBM = 256, kMapBK = 8, kMapStages = 4, kMapThreads = 128, kMapN = 32, kMapCtasPerSm = 3;

__device__ __forceinline__ void map_cp16(uint32_t dst, const void *src, bool pred) {
    const int sz = pred ? 16 : 0;                                   // src-size 0: zero fill (pixel tail of a view)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

// Persistent CTAs (3 per SM): the weights are staged once per CTA and the cp.async ring runs across tile boundaries,
// so the pipeline never drains between the ~3 tiles a CTA owns.
template <typename T>
__global__ void __launch_bounds__(kMapThreads, kMapCtasPerSm)
k_map_features(const T *__restrict__ feat, int64_t sv, int64_t sc, int n_pix, int channels, const float *__restrict__ weight,
               const float *__restrict__ bias, float *__restrict__ out, int tiles_per_view, int n_tiles) {
    extern __shared__ __align__(16) float smem[];
    float *sW = smem;                                               // [channels][32]
    float *sA = smem + (size_t)channels * kMapN;                    // [stages][8][256]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sA_u = (uint32_t)__cvta_generic_to_shared(sA);
    constexpr int kElt = 16 / (int)sizeof(T);                       // elements per 16-byte chunk
    constexpr int kChunksPerRow = kMapBM / kElt;
    const int n_k = channels / kMapBK;
    const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_tiles * n_k;                               // (tile, channel block) steps of this CTA

    auto load_step = [&](int step) {                                // step -> (tile, kb); slot = step % stages
        const int t = step / n_k, kb = step - t * n_k;
        const int tile_id = (int)blockIdx.x + t * (int)gridDim.x;
        const int v = tile_id / tiles_per_view, p0 = (tile_id - v * tiles_per_view) * kMapBM;
        const T *src = feat + (int64_t)v * sv + p0;
        const int s = step % kMapStages;
        for (int i = tid; i < kMapBK * kChunksPerRow; i += kMapThreads) {
            const int r = i / kChunksPerRow, ch = i - r * kChunksPerRow;
            const bool ok = p0 + ch * kElt < n_pix;                // n_pix is a multiple of the chunk (checked by the host)
            map_cp16(sA_u + (uint32_t)(((s * kMapBK + r) * kMapBM) * sizeof(T) + ch * 16),
                     src + (int64_t)(kb * kMapBK + r) * sc + (ok ? ch * kElt : 0), ok);
        }
    };
    for (int st = 0; st < kMapStages - 1; ++st) {
        if (st < total) load_step(st);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int i = tid; i < channels * kMapN; i += kMapThreads) {     // Wt[c][j] = weight[j][c]; conflict-free stores
        const int c = i / kMapN, j = i - c * kMapN;
        sW[i] = __ldg(weight + (size_t)j * channels + c);
    }
    float bv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bv[j] = bias != nullptr ? __ldg(bias + warp * 8 + j) : 0.0f;
    float acc[8][8];
    int step = 0;
    for (int t = 0; t < my_tiles; ++t) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = bv[j];
        for (int kb = 0; kb < n_k; ++kb, ++step) {
            asm volatile("cp.async.wait_group %0;" ::"n"(kMapStages - 2) : "memory");
            __syncthreads();                                        // step landed; the slot of step - 1 is free for reuse
            if (step + kMapStages - 1 < total) load_step(step + kMapStages - 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
            const T *a = reinterpret_cast<const T *>(sA) + (size_t)(step % kMapStages) * kMapBK * kMapBM;
            const float *w = sW + (size_t)kb * kMapBK * kMapN + warp * 8;
#pragma unroll
            for (int k = 0; k < kMapBK; ++k) {
                float av[8];
                if constexpr (sizeof(T) == 4) {
                    const float4 a0 = *reinterpret_cast<const float4 *>(a + k * kMapBM + 4 * lane);
                    const float4 a1 = *reinterpret_cast<const float4 *>(a + k * kMapBM + 128 + 4 * lane);
                    av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
                    av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        av[i] = to_f32<T>(a[k * kMapBM + 4 * lane + i]);
                        av[4 + i] = to_f32<T>(a[k * kMapBM + 128 + 4 * lane + i]);
                    }
                }
                const float4 w0 = *reinterpret_cast<const float4 *>(w + k * kMapN);
                const float4 w1 = *reinterpret_cast<const float4 *>(w + k * kMapN + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
            }
        }
        // channels-last rows: pixel p -> out[(v * n_pix + p) * 32 + 8 warp .. + 8): two float4 per pixel
        const int tile_id = (int)blockIdx.x + t * (int)gridDim.x;
        const int v = tile_id / tiles_per_view, p0 = (tile_id - v * tiles_per_view) * kMapBM;
        float *dst = out + ((int64_t)v * n_pix + p0) * kMapN + warp * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = (i < 4 ? 4 * lane + i : 128 + 4 * lane + (i - 4));
            if (p0 + p < n_pix) {
                float4 *o = reinterpret_cast<float4 *>(dst + (int64_t)p * kMapN);
                __stcs(o, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                __stcs(o + 1, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
            }
        }
    }
}

}  // namespace nd

using namespace nd;

extern "C" int nd_map_features(const nd_maps *features, const float *weight, const float *bias, int out_channels,
                               float *mapped, void *stream) {
    ND_REQUIRE(features && features->data && weight && mapped, ND_ERR_BAD_ARG, "nd_map_features: null pointer");
    ND_REQUIRE(out_channels == kMapN, ND_ERR_BAD_SHAPE, "nd_map_features: %d output channels (this build maps to 32)", out_channels);
    const int elt = features->dtype == ND_F32 ? 4 : 2;
    const int64_t n_pix = (int64_t)features->height * features->width;
    ND_REQUIRE(features->n_views > 0 && features->channels > 0 && features->channels % kMapBK == 0 &&
                   features->channels <= 1024,
               ND_ERR_BAD_SHAPE, "nd_map_features: channel count %d must be a multiple of 8 (<= 1024)", features->channels);
    ND_REQUIRE(features->stride_x == 1 && features->stride_y == features->width, ND_ERR_BAD_SHAPE,
               "nd_map_features: planes must be contiguous");
    ND_REQUIRE((n_pix * elt) % 16 == 0 && (features->stride_c * elt) % 16 == 0 && (features->stride_v * elt) % 16 == 0 &&
                   (reinterpret_cast<uintptr_t>(features->data) & 15) == 0 && (reinterpret_cast<uintptr_t>(mapped) & 15) == 0,
               ND_ERR_BAD_ALIGNMENT, "nd_map_features: planes and output must be 16-byte aligned");
    const int tiles = (int)ceil_div(n_pix, kMapBM);
    const size_t smem = ((size_t)features->channels * kMapN) * sizeof(float) + (size_t)kMapStages * kMapBK * kMapBM * elt;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = features->n_views * tiles;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(n_tiles < sms * kMapCtasPerSm ? n_tiles : sms * kMapCtasPerSm);
    cudaError_t e;
    if (features->dtype == ND_F32) {
        auto kern = k_map_features<float>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            kern<<<grid, kMapThreads, smem, st>>>((const float *)features->data, features->stride_v, features->stride_c,
                                                  (int)n_pix, features->channels, weight, bias, mapped, tiles, n_tiles);
    } else {
        auto kern = k_map_features<__nv_bfloat16>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            kern<<<grid, kMapThreads, smem, st>>>((const __nv_bfloat16 *)features->data, features->stride_v,
                                                  features->stride_c, (int)n_pix, features->channels, weight, bias, mapped,
                                                  tiles, n_tiles);
    }
    if (e != cudaSuccess) {
        set_error("nd_map_features: %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    ND_CUDA_LAUNCH_CHECK("k_map_features");
    return ND_OK;
}
