// fp64 DFMA tile configurations of mtm_simt_kernel (see mtm_simt.cuh).
#include "mtm_simt_dispatch.cuh"

namespace b200 {

namespace {
const TileConfig kCfg[] = {
    {"dfma_128x64x8_t8x4", 128, 64, 8, 256, 2},
    {"dfma_64x64x8_t4x4", 64, 64, 8, 256, 3},
    {"dfma_64x128x8_t4x8", 64, 128, 8, 256, 2},
    {"dfma_128x128x8_t8x8", 128, 128, 8, 256, 1},
};
}  // namespace

int simt_f64_num_configs() { return (int)(sizeof(kCfg) / sizeof(kCfg[0])); }
const TileConfig& simt_f64_config(int cfg) { return kCfg[cfg]; }

cudaError_t launch_simt_f64(int cfg, double* C, const double* A, const double* B, const MtmShape& s,
                            int amode, int bmode, int vec_c, cudaStream_t stream) {
#define D0(AM, BMD) mtm_simt_kernel<double, 128, 64, 8, 8, 4, 2, AM, BMD>
#define D1(AM, BMD) mtm_simt_kernel<double, 64, 64, 8, 4, 4, 3, AM, BMD>
#define D2(AM, BMD) mtm_simt_kernel<double, 64, 128, 8, 4, 8, 2, AM, BMD>
#define D3(AM, BMD) mtm_simt_kernel<double, 128, 128, 8, 8, 8, 1, AM, BMD>
    switch (cfg) {
        case 0: B200_DISPATCH_MODES(D0, 128, 64, 256);
        case 1: B200_DISPATCH_MODES(D1, 64, 64, 256);
        case 2: B200_DISPATCH_MODES(D2, 64, 128, 256);
        case 3: B200_DISPATCH_MODES(D3, 128, 128, 256);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace b200
