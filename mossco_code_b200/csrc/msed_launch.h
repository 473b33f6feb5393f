// msed_launch.h -- launchers of the stepping kernels.  Each family of kernel templates is instantiated in a
// translation unit of its own (msed_tu_*.cu), so that the library builds in parallel; msed.cu (host side,
// C ABI, controllers, helper kernels) only sees these functions.
#pragma once

#include "msed_types.cuh"

namespace msed {

// column_kernel<MODEL, OP, PROFILE3, STREAM_POR> (msed_column.cuh): one RHS evaluation / attempt / RK stage
cudaError_t tu_launch_column(int model, bool profile3, bool stream_por, int op, const KParams &p, cudaStream_t s);

// pair_kernel (msed_pair.cuh): two accepted sub-steps per launch.  DENIT <- p.denit_out, COLMAP <- p.colmap,
// OVR <- p.in_ovr; the caller has already converted col0/col_end to a range of the wet-column list
cudaError_t tu_launch_pair(int model, bool adaptive, const KParams &p, cudaStream_t s);
cudaError_t tu_enable_pair_smem();
constexpr int TU_PAIR_CTAS_PER_SM = 384 / COL_BLOCK;   // resident pair_kernel CTAs per SM (msed_pair.cuh, PAIR_MIN_BLOCKS)

// chain_kernel (msed_chain.cuh): nsteps ode_solver calls per launch, warp per column, knum <= 64
cudaError_t tu_launch_chain(int model, bool adaptive, bool clip, const KParams &p, int nsteps, cudaStream_t s);
// rk_chain_kernel (msed_chain.cuh): the same for RK4 / RK4-3/8, four stages per call on the column in registers
cudaError_t tu_launch_rk_chain(int model, int method, bool clip, const KParams &p, int m, cudaStream_t s);
constexpr int TU_RK_CHAIN_MAX_STEPS = 4;  // ode_solver calls per launch (16 RHS evaluations, as an Euler chain)
constexpr int TU_CHAIN_MAX_LAYERS = 64;   // one layer per lane up to 32, two above
constexpr int TU_SPINUP_MAX_LAYERS = 64;  // spinup_kernel: one layer per lane up to 32, two above
constexpr int TU_CHAIN_MAX_STEPS = 16;   // steps per launch: bounds the work a failed speculation throws away

// rk_pair_kernel (msed_rkpair.cuh): which = 0 for stages 1+2, 1 for stages 3+4
cudaError_t tu_launch_rk_pair(int model, int method, int which, const KParams &p, cudaStream_t s);
cudaError_t tu_enable_rk_smem();

// rk_quad_kernel (msed_rkquad.cuh): the four stages of a Runge-Kutta call in one launch, thread per column
cudaError_t tu_launch_rk_quad(int model, int method, const KParams &p, cudaStream_t s);
cudaError_t tu_enable_rk_quad_smem();
constexpr int TU_RK_QUAD_MIN_LAYERS = 5;

// spinup_kernel (msed_spinup.cuh): the 1-D pre-simulation of a batch of members, warp per member, knum <= 64
cudaError_t tu_launch_spinup(int model, const KParams &p, const SpinupArgs &a, cudaStream_t s);

}  // namespace msed
