// fp64 DMMA (mma.sync m8n8k4) tile configurations of mtm_dmma_kernel (see mtm_simt.cuh).
#include "mtm_simt_dispatch.cuh"

namespace b200 {

namespace {
const TileConfig kCfg[] = {
    {"dmma_128x128x8_w4x4", 128, 128, 8, 512, 1},
    {"dmma_64x64x8_w2x2", 64, 64, 8, 128, 4},
    {"dmma_128x64x8_w4x2", 128, 64, 8, 256, 2},
    {"dmma_64x128x8_w2x4", 64, 128, 8, 256, 2},
    {"dmma_64x64x16_w2x2", 64, 64, 16, 128, 4},     // deeper K slices: half the barriers per FMA
};
}  // namespace

int dmma_f64_num_configs() { return (int)(sizeof(kCfg) / sizeof(kCfg[0])); }
const TileConfig& dmma_f64_config(int cfg) { return kCfg[cfg]; }

cudaError_t launch_dmma_f64(int cfg, double* C, const double* A, const double* B, const MtmShape& s,
                            int amode, int bmode, int vec_c, cudaStream_t stream) {
#define M0(AM, BMD) mtm_dmma_kernel<128, 128, 8, 4, 4, 1, AM, BMD>
#define M1(AM, BMD) mtm_dmma_kernel<64, 64, 8, 2, 2, 4, AM, BMD>
#define M2(AM, BMD) mtm_dmma_kernel<128, 64, 8, 4, 2, 2, AM, BMD>
#define M3(AM, BMD) mtm_dmma_kernel<64, 128, 8, 2, 4, 2, AM, BMD>
#define M4(AM, BMD) mtm_dmma_kernel<64, 64, 16, 2, 2, 4, AM, BMD>
    switch (cfg) {
        case 0: B200_DISPATCH_MODES(M0, 128, 128, 512);
        case 1: B200_DISPATCH_MODES(M1, 64, 64, 128);
        case 2: B200_DISPATCH_MODES(M2, 128, 64, 256);
        case 3: B200_DISPATCH_MODES(M3, 64, 128, 256);
        case 4: B200_DISPATCH_MODES(M4, 64, 64, 128);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace b200
