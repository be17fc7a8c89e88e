// Internal launch interface between the C-ABI layer (mtm_api.cu) and the kernel translation units.
#pragma once

#include "mtm_common.cuh"

namespace b200 {

struct TileConfig {
    const char* name;
    int bm, bn, bk;
    int threads;
    int min_blocks;  // __launch_bounds__ residency the kernel was compiled for
};

// CUDA-core kernels (mtm_simt_f32.cu / mtm_simt_f64.cu)
int simt_f32_num_configs();
const TileConfig& simt_f32_config(int cfg);
cudaError_t launch_simt_f32(int cfg, float* C, const float* A, const float* B, const MtmShape& s,
                            int amode, int bmode, int vec_c, cudaStream_t stream);

int simt_f64_num_configs();
const TileConfig& simt_f64_config(int cfg);
cudaError_t launch_simt_f64(int cfg, double* C, const double* A, const double* B, const MtmShape& s,
                            int amode, int bmode, int vec_c, cudaStream_t stream);

// TMA-fed fp32 FFMA kernel (mtm_ffma_tma.cu).  Appears to callers as SIMT configs
// simt_f32_num_configs() ... + ffma_tma_num_configs() - 1.  `ws` holds packed operand planes.
int ffma_tma_num_configs();
const TileConfig& ffma_tma_config(int cfg);
size_t ffma_tma_workspace_bytes(const MtmShape& s);
// reuse_b != 0: the re-laid B operand left in `ws` by the previous call (same B, same K and N) is
// still valid — skip its pack/split pass (used by the slab pipeline of the host-pointer entry).
cudaError_t launch_ffma_tma_f32(int cfg, float* C, const float* A, const float* B, const MtmShape& s, void* ws,
                                size_t ws_bytes, int vec_c, int reuse_b, cudaStream_t stream, int* launches);

// fp64 tensor-core kernels (mtm_dmma_f64.cu)
int dmma_f64_num_configs();
const TileConfig& dmma_f64_config(int cfg);
cudaError_t launch_dmma_f64(int cfg, double* C, const double* A, const double* B, const MtmShape& s,
                            int amode, int bmode, int vec_c, cudaStream_t stream);

// TMA-fed fp64 DMMA kernel (mtm_dmma_tma.cu).  Appears to callers as DMMA configs
// dmma_f64_num_configs() ... + dmma_tma_num_configs() - 1.  `ws` holds K-contiguous operand planes.
int dmma_tma_num_configs();
const TileConfig& dmma_tma_config(int cfg);
size_t dmma_tma_workspace_bytes(const MtmShape& s);
cudaError_t launch_dmma_tma_f64(int cfg, double* C, const double* A, const double* B, const MtmShape& s, void* ws,
                                size_t ws_bytes, int vec_c, int reuse_b, cudaStream_t stream, int* launches);

// fp32 3xTF32 tcgen05 path (mtm_tf32.cu).  `ws` is device workspace of tf32_workspace_bytes(): the lo planes
// (and, for operands the TMA cannot fetch in place, gathered hi planes).  a_mode / b_mode report how each
// operand was fed: 0 = in place, K-major; 1 = in place, MN-major; 2 = packed (tf32_operand_mode_name).
size_t tf32_workspace_bytes(const MtmShape& s, const float* A, const float* B);
// split_k: 0 = automatic (tf32_auto_split), else the number of K splits per output tile (added into C in order).
cudaError_t launch_3xtf32_f32(float* C, const float* A, const float* B, const MtmShape& s, void* ws,
                              size_t ws_bytes, int cfg, int reuse_b, int reserve_sms, int split_k, cudaStream_t stream,
                              int* launches, int* a_mode = nullptr, int* b_mode = nullptr, int* split_used = nullptr,
                              int* cfg_used = nullptr);
int tf32_auto_split(int64_t tiles, int nkb, int slots);
int tf32_tail_split(int64_t tiles, int nkb, int slots);
double tf32_wave_efficiency(int64_t tiles, int nkb, int slots);
const char* tf32_operand_mode_name(int mode);
int tf32_num_configs();
cudaError_t tf32_preload_kernels();   // force-load every kernel of the path (multi-GPU drivers wait in-kernel)
const TileConfig& tf32_config(int cfg);

// Matrix-times-vector (mtv.cu): c[i] (op)= sum_k a[i*s_i + k*s_k] * b[k]; `ws` holds chunk partials.
size_t mtv_workspace_bytes(int64_t M, int elem_size, int sm_count);
cudaError_t launch_mtv_f32(float* c, const float* a, int64_t M, int64_t K, int64_t s_i, int64_t s_k, const float* b,
                           int accumulate, void* ws, size_t ws_bytes, int sm_count, cudaStream_t stream, int* launches,
                           const char** name);
cudaError_t launch_mtv_f64(double* c, const double* a, int64_t M, int64_t K, int64_t s_i, int64_t s_k, const double* b,
                           int accumulate, void* ws, size_t ws_bytes, int sm_count, cudaStream_t stream, int* launches,
                           const char** name);

// Transpose (trans.cu): c(j,i) = a(i,j) with a[i*sa_i + j*sa_j], c[j*sc_j + i*sc_i]; in place: n x n, (i,j) at a[i + j*n].
cudaError_t launch_transpose_f32(float* c, const float* a, int64_t M, int64_t N, int64_t sa_i, int64_t sa_j,
                                 int64_t sc_j, int64_t sc_i, cudaStream_t stream);
cudaError_t launch_transpose_f64(double* c, const double* a, int64_t M, int64_t N, int64_t sa_i, int64_t sa_j,
                                 int64_t sc_j, int64_t sc_i, cudaStream_t stream);
cudaError_t launch_transpose_inplace_f32(float* a, int64_t n, cudaStream_t stream);
cudaError_t launch_transpose_inplace_f64(double* a, int64_t n, cudaStream_t stream);

// Operand replication over NVLink / NVSwitch multicast (replicate.cu).
cudaError_t launch_replicate_push(void* const* dst, int n_dst, int multicast, const void* src, size_t bytes,
                                  uint32_t* const* flag_dst, int n_flag_dst, int flag_multicast, uint32_t flag_value,
                                  int ctas, cudaStream_t stream);
cudaError_t launch_replicate_push_2d(void* const* dst, int n_dst, int multicast, const void* src, size_t rows,
                                     size_t row_bytes, size_t src_pitch, size_t dst_pitch, uint32_t* const* flag_dst,
                                     int n_flag_dst, int flag_multicast, uint32_t flag_value, int ctas, cudaStream_t stream);
cudaError_t replicate_preload_kernels();
cudaError_t launch_flag_wait(const uint32_t* flag, uint32_t value, int count, int stride, int skip, cudaStream_t stream);
cudaError_t launch_flag_signal(uint32_t* flag, uint32_t value, cudaStream_t stream);

}  // namespace b200
