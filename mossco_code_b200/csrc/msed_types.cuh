// msed_types.cuh -- shared device-side types of the fabm_sediment column solver: the control block of the
// step loop, the kernel parameter block, integrator stage ids.  Included by every translation unit
// (msed.cu and the msed_tu_*.cu files that instantiate the stepping kernels).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda.h>   // CUtensorMap only: the driver entry point is looked up at run time, nothing links libcuda
#include <cuda_runtime.h>
#include <type_traits>

#include "../../include/msed.h"

namespace msed {

constexpr int NV = MSED_NVAR;
constexpr int MAXK = MSED_MAX_LAYERS;
constexpr int NPART = 3;  // ldetC, sdetC, detP are particulate (main.F90:92-101)
// tunables of the column kernel (overridable with -D for the sweeps in tools/tune_sweep.sh)
#ifndef MSED_COL_BLOCK
#define MSED_COL_BLOCK 128
#endif
#ifndef MSED_COL_MIN_BLOCKS
#define MSED_COL_MIN_BLOCKS 4
#endif
#ifndef MSED_RING_STAGES
#define MSED_RING_STAGES 4
#endif
constexpr int COL_BLOCK = MSED_COL_BLOCK;            // threads (= columns) per CTA of the column kernel
constexpr int COL_MIN_BLOCKS = MSED_COL_MIN_BLOCKS;  // 4 CTAs/SM -> <=128 registers/thread, 16 warps/SM
#define MSED_NFLAGS 40
constexpr int FLAG_FAIL = 4;                         // chain_kernel: 64 - (step of the launch a rejectable violation stopped a warp at)
constexpr int FLAG_UP0 = 8;                          // first "planned rejection seen" slot of Ctl::flags
constexpr int MAX_UP_SLOTS = MSED_NFLAGS - FLAG_UP0; // planned rejections one fused group can hold
constexpr int MAX_PLAN_DEPTH = 2;                    // fused launches cover steps that run at dt, dt/4 or dt/16

// integrator stage executed by the column kernel
enum Op : int {
    OP_RHS = 0,       // get_rhs only
    OP_EULER,         // solver_library.F90:99-102
    OP_ADAPTIVE,      // one attempt of :104-140
    OP_RK4_S1, OP_RK4_S2, OP_RK4_S3, OP_RK4_S4,         // :142-163
    OP_RK38_S1, OP_RK38_S2, OP_RK38_S3, OP_RK38_S4      // :164-185
};

// device-resident control block of the step loop (one per handle)
struct Ctl {
    double dt;          // requested ode_solver dt
    double dt_int;      // integrated time inside the current ode_solver call (:106)
    double dt_red;      // current reduced sub-step (:107,:127)
    double dt_min;      // type_rhs_driver%dt_min
    double last_min_dt; // :44
    long long steps_done, steps_target;
    long long rhs_evals, subcycles;
    long long fused_steps, fused_launches;   // ode_solver calls / launches committed by plan_controller_kernel
    int cur;            // which of buf[0..1] is sed%conc
    int flags[MSED_NFLAGS];  // [0] relative-change violation (:121)  [1] NaN (component :2392);
                        // [2],[3] the same for the second stage of a fused pair (msed_pair.cuh);
                        // [FLAG_UP0 + s]: a fused launch that PLANS a rejected attempt (sub-cycling, :126-128) saw
                        // the violation that rejects it -- slot s = (step of the group, level); all planned slots
                        // must be up for the group to be committed (msed.cu run_steps)
    int pairs_disabled; // a fused pair could not be committed: fall back to single steps
    int pair_failures;
    int nan_detected;
    int stop;
    int do_clip;        // component wrapper (check_NaN + clip) on/off
    int diagnostics;    // adaptive_solver_diagnostics
    int minloc_request; // set when last_min_dt decreased (:131-135)
    int step_completed; // 1 if the last controller invocation finished an ode_solver call
    // accept/reject history of the ode_solver call in progress and of the last completed one: what the next
    // call's fused plan is predicted from (msed.cu run_steps)
    int step_rej_first; // attempts rejected before the first accepted sub-step of the current call (:126-128)
    int step_rej_later; // ... after it
    int step_accepts;   // accepted sub-steps of the current call
    int last_depth;     // step_rej_first of the last completed call: it ran at dt/4^last_depth
    int last_irregular; // ... and whether it rejected anything after its first accepted sub-step
    int fail_step;      // set when a chain launch is not committed: a step of the launch (0-based) at or before which it
                        // went wrong -- the steps in front of it can be re-run as a shorter chain; -1 unknown
};

struct OmexDev {  // hzg_omexdia_p parameters, rates already per second
    double rLabile, rSemilabile, NCrLdet, NCrSdet, PAds_rS, PAdsODU, rNH3Ads, CprodMax;
    double rnit, ksO2nitri, rODUox, ksO2oduox, ksO2oxic, ksNO3denit, kinO2denit;
    double kinNO3anox, kinO2anox, E_a;
    double minimum[NV];
};

struct KParams {
    double *buf[2];          // ping-pong state, [nvar][K][ld]
    double *aux1, *aux2;     // RK accumulators
    double *rhs_out;         // OP_RHS target
    const double *por;       // [K][ld]
    const double *bdys;      // [nvar+1][ld]
    double *fluxes;          // [nvar][ld]
    const unsigned char *mask;  // [ncol]
    Ctl *ctl;
    size_t ld;
    int ncol, K, inum;
    int col0, col_end;       // column range [col0, col_end) of this launch (chunked launches overlap PCIe)
    int i_offset, j_offset;
    int bcup_diss, bcup_part, profile;
    int use_ctl;             // 0: OP_RHS / plain launch with p.dt and buffer 0
    int por_mode;            // 0: 3-D porosity field, 1: portab[k], 2: por(:,:,1)*portab[k]
    double dt;
    double fac;              // 1 + relative_change_min
    double bioturbation, diffusivity;
    double pom_flux_rate;    // pom_flux_max/86400
    double beta, b, L1, L2, poc_factor[2], cumdepth_last;
    OmexDev om;
    double dz[MAXK], rdzc[MAXK], bf[MAXK], e1[MAXK], e2[MAXK], portab[MAXK];
    double *denit_out;       // [K][ld]: FABM denit diagnostic of the second step of a call's last pair, or null
    const int *colmap;       // pair_kernel on a masked tile: indices of the wet columns, ascending; col0/col_end
                             // then count wet columns (null: identity)
    int min_zero;            // every state variable's minimum is +0: pair_kernel may clip lazily (msed_pair.cuh)
    int feed_bulk;           // pair_kernel: the tensor maps below are valid and the tile has no masked column, so the
                             // state may be fed to shared memory by TMA box copies (msed_pair.cuh, FEED_BULK)
    // TMA descriptors of the state buffers as 3-D tensors [nvar][K][ld] of fp64, box = 32 columns x 1 layer x
    // 8 variables: of buf[0], buf[1] and of in_ovr
    alignas(64) CUtensorMap tmap[3];
    const double *in_ovr;    // pair_kernel<.., OVR>: explicit input / output state buffers of a launch inside a
    double *out_ovr;         // chunk-major sequence (msed.cu run_steps), instead of buf[cur] / buf[1-cur]
    // ---- the plan of a fused launch (msed.cu run_steps): which accepted sub-steps of the reference's attempt
    // ---- sequence (solver_library.F90:104-140) the launch performs -----------------------------------------
    long long gate_steps;    // Ctl::steps_done the plan was made for; a launch that finds another value does nothing
    double dt_acc;           // length of every accepted sub-step of the launch: dt/4^depth
    int depth;               // rejected attempts in front of the first sub-step of a step (at dt_acc*4^depth .. dt_acc*4)
    int pair_kind;           // pair_kernel: PAIR_FULL / PAIR_FIRST / PAIR_MID / PAIR_LAST
    int up_slot;             // Ctl::flags slot of the first planned rejection this launch has to see
};

// What the two stages of a pair_kernel launch are, in terms of the sub-steps of one ode_solver call:
enum PairKind : int {
    PAIR_FULL = 0,   // two whole steps (each accepted at dt on the first attempt): check_NaN + clip after both
    PAIR_FIRST,      // the first two sub-steps of a step; stage A also tests the `depth` larger step sizes that
                     // must be rejected for the step to run at dt_acc
    PAIR_MID,        // two inner sub-steps
    PAIR_LAST        // the last two sub-steps of a step: check_NaN + clip after stage B
};

// how a committed fused group changes the control block: computed on the host from the plan, applied by
// plan_controller_kernel when every flag agrees with the plan
struct PlanCommit {
    long long gate_steps;   // Ctl::steps_done at the start of the group
    long long steps;        // ode_solver calls completed by the group
    long long rhs_evals;    // attempts: accepted sub-steps + planned rejections
    long long subcycles;    // planned rejections (:127-128)
    int up_slots;           // flags[FLAG_UP0 .. FLAG_UP0+up_slots) must all be up
    int own_rejectable;     // a violation at dt_acc would be rejected (dt_acc > dt_min, :126): must not be up
    int flip;               // state ends in the other buffer
    int launches;           // fused launches (per chunk) the group consists of
    int depth;              // Ctl::last_depth after the group
    double dt_int, dt_red;  // :106-107 after the group (0 and dt when it ends on a step boundary)
};



// arguments of spinup_kernel (msed_spinup.cuh) beside the kernel parameter block
struct SpinupArgs {
    const OmexDev *om;       // [nmember] reaction parameters, or null: KParams::om for every member
    double *last_min_dt;     // [nmember] out
    int *grid_cell;          // [nmember][4] out (1-based i, j, k, n; -99 when never set)
    long long *counters;     // [nmember][2] out: rhs evaluations, sub-cycle warnings
    long long nsteps;
    double dt;               // dt_spinup
    double dt_min;
    double last_min_dt0;     // solver_library.F90:44
    int method;
};

}  // namespace msed
