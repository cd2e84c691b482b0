// A C++ host that binds the C ABI directly (no Python, no torch): the geometry plan of one scene is built once and
// reused for every feature stack lifted with the same cameras -- what nerfdet.py:155-181 does per scene.
//
//   g++ -std=c++17 -Iinclude -I/usr/local/cuda/include examples/host_plan_lift.cpp
//       -Lnerfdet_b200/lib -lnerfdet_lift -L/usr/local/cuda/lib64 -lcudart -o host_plan_lift
// (compiled and linked by tests/test_cabi_symbols.py on the CPU tier; it is documentation of the call sequence that
// nerfdet_b200/ops.py:LiftPlan makes through ctypes, which is what the GPU tests exercise)
//
// Inputs here are zeros / an identity-like camera: the point of the file is the call sequence, the ownership rules
// (every buffer is the caller's, the library never allocates or synchronises) and the error handling.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "nerfdet_lift.h"

#define CUDA_OK(call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            std::fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));               \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)
#define ND_OK_OR_DIE(call)                                                                 \
    do {                                                                                   \
        if ((call) != ND_OK) {                                                             \
            std::fprintf(stderr, "%s: %s\n", #call, nd_last_error_string());               \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

int main() {
    // nerfdet_res50_2x_low_res: 50 views of [256, 60, 80] fp32 stride-4 features used as the [:, :, :59, :80] slice
    const int nv = 50, C = 256, Hp = 60, Wp = 80, H = 59, W = 80;
    const int gx = 40, gy = 40, gz = 16;
    const int64_t N = (int64_t)gx * gy * gz;
    std::printf("libnerfdet_lift ABI version %d\n", nd_version());

    float *features = nullptr, *points = nullptr, *projection = nullptr, *mean = nullptr, *cov = nullptr;
    int64_t *count = nullptr;
    CUDA_OK(cudaMalloc(&features, sizeof(float) * nv * C * Hp * Wp));
    CUDA_OK(cudaMemset(features, 0, sizeof(float) * nv * C * Hp * Wp));
    CUDA_OK(cudaMalloc(&points, sizeof(float) * 3 * N));            // get_points(): [3][X][Y][Z], Z fastest
    CUDA_OK(cudaMalloc(&projection, sizeof(float) * nv * 12));      // _compute_projection(): [nv][3][4]
    CUDA_OK(cudaMalloc(&mean, sizeof(float) * C * N));
    CUDA_OK(cudaMalloc(&cov, sizeof(float) * C * N));
    CUDA_OK(cudaMalloc(&count, sizeof(int64_t) * N));
    std::vector<float> h_pts(3 * N), h_proj(nv * 12, 0.0f);
    for (int64_t n = 0; n < N; ++n) {
        const int iz = (int)(n % gz), iy = (int)((n / gz) % gy), ix = (int)(n / ((int64_t)gz * gy));
        h_pts[n] = (ix - gx / 2 + 0.5f) * 0.16f;
        h_pts[N + n] = (iy - gy / 2 + 0.5f) * 0.16f;
        h_pts[2 * N + n] = (iz + 0.5f) * 0.2f + 2.0f;
    }
    for (int v = 0; v < nv; ++v) {                                  // a pinhole looking down +z, shifted per view
        float *p = h_proj.data() + v * 12;
        p[0] = 70.0f; p[2] = 40.0f + 0.1f * v; p[5] = 70.0f; p[6] = 29.5f; p[10] = 1.0f;
    }
    CUDA_OK(cudaMemcpy(points, h_pts.data(), sizeof(float) * 3 * N, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(projection, h_proj.data(), sizeof(float) * nv * 12, cudaMemcpyHostToDevice));

    cudaStream_t stream;
    CUDA_OK(cudaStreamCreate(&stream));

    // the reference's non-contiguous slice is described by strides, in elements
    nd_maps maps{};
    maps.data = features;
    maps.dtype = ND_F32;
    maps.n_views = nv; maps.channels = C; maps.height = H; maps.width = W;
    maps.stride_v = (int64_t)C * Hp * Wp; maps.stride_c = (int64_t)Hp * Wp; maps.stride_y = Wp; maps.stride_x = 1;

    nd_lift_options opt{};                                          // zeros = automatic everything
    opt.grid_x = gx; opt.grid_y = gy; opt.grid_z = gz;              // lets the plan use compact 4 x 8 x 4 quads

    const size_t plan_bytes = nd_lift_plan_bytes(&maps, N, &opt);
    if (plan_bytes == 0) {                                          // layout the plane-resident kernel does not take
        void *ws = nullptr;
        const size_t ws_bytes = nd_lift_workspace_bytes(&maps, N, &opt);
        CUDA_OK(cudaMalloc(&ws, ws_bytes));
        ND_OK_OR_DIE(nd_lift_mean_var(&maps, points, projection, N, nullptr, mean, cov, count, ws, ws_bytes, &opt, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        CUDA_OK(cudaFree(ws));
    } else {
        void *plan = nullptr;
        CUDA_OK(cudaMalloc(&plan, plan_bytes));
        // once per scene (points / projection / depth): three small kernels
        ND_OK_OR_DIE(nd_lift_plan_build(&maps, points, projection, N, /*depth_resized=*/nullptr, /*voxel_z=*/0.0f, plan,
                                        plan_bytes, &opt, stream));
        // every lift with these cameras: one kernel; the caller counts its launches on this plan
        for (uint32_t launch = 0; launch < 3; ++launch)
            ND_OK_OR_DIE(nd_lift_plan_mean_var(&maps, plan, plan_bytes, N, launch, /*n_views_total=*/0, /*alpha=*/nullptr,
                                               mean, cov, count, &opt, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        CUDA_OK(cudaFree(plan));
    }

    std::vector<int64_t> h_count(N);
    CUDA_OK(cudaMemcpy(h_count.data(), count, sizeof(int64_t) * N, cudaMemcpyDeviceToHost));
    int64_t seen = 0;
    for (int64_t c : h_count) seen += c;
    std::printf("%lld valid voxel-views of %lld\n", (long long)seen, (long long)(N * nv));

    CUDA_OK(cudaStreamDestroy(stream));
    for (void *p : {(void *)features, (void *)points, (void *)projection, (void *)mean, (void *)cov, (void *)count})
        CUDA_OK(cudaFree(p));
    return 0;
}
