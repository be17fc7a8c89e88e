// fp32 FFMA tile configurations of mtm_simt_kernel (see mtm_simt.cuh).
#include "mtm_simt_dispatch.cuh"

namespace b200 {

namespace {
//                         name          BM   BN  BK  threads  minBlocks
const TileConfig kCfg[] = {
    {"ffma_128x128x8_t8x8", 128, 128, 8, 256, 2},
    {"ffma_64x64x16_t4x4", 64, 64, 16, 256, 3},
    {"ffma_128x128x16_t8x8", 128, 128, 16, 256, 2},
    {"ffma_256x128x8_t16x8", 256, 128, 8, 256, 1},
    {"ffma_128x256x8_t8x16", 128, 256, 8, 256, 1},
};
}  // namespace

int simt_f32_num_configs() { return (int)(sizeof(kCfg) / sizeof(kCfg[0])); }
const TileConfig& simt_f32_config(int cfg) { return kCfg[cfg]; }

cudaError_t launch_simt_f32(int cfg, float* C, const float* A, const float* B, const MtmShape& s,
                            int amode, int bmode, int vec_c, cudaStream_t stream) {
#define K0(AM, BMD) mtm_simt_kernel<float, 128, 128, 8, 8, 8, 2, AM, BMD>
#define K1(AM, BMD) mtm_simt_kernel<float, 64, 64, 16, 4, 4, 3, AM, BMD>
#define K2(AM, BMD) mtm_simt_kernel<float, 128, 128, 16, 8, 8, 2, AM, BMD>
#define K3(AM, BMD) mtm_simt_kernel<float, 256, 128, 8, 16, 8, 1, AM, BMD>
#define K4(AM, BMD) mtm_simt_kernel<float, 128, 256, 8, 8, 16, 1, AM, BMD>
    switch (cfg) {
        case 0: B200_DISPATCH_MODES(K0, 128, 128, 256);
        case 1: B200_DISPATCH_MODES(K1, 64, 64, 256);
        case 2: B200_DISPATCH_MODES(K2, 128, 128, 256);
        case 3: B200_DISPATCH_MODES(K3, 256, 128, 256);
        case 4: B200_DISPATCH_MODES(K4, 128, 256, 256);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace b200
