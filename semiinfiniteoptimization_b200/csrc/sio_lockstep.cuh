// sio_lockstep.cuh -- state and launchers of the device-resident lock-step pose solver (sio_lockstep.cu)
#pragma once
#include "sio_common.cuh"

namespace sio {

struct LsSettings {
    int32_t max_iters, cap;
    double xeps2, inflation, drop_value, min_cv, w0, w_min, excl, reg;
};

struct LsState {
    int64_t B, N;
    int32_t cap;
    const GridDev* grid;
    const double* cloud_world;        // [N*3] world coordinates of the environment cloud, caller order
    double Tenv[12];
    double *R, *t, *Rdes, *tdes;      // [B*9], [B*3]
    double* P; long long* key; int32_t* cnt;
    double *D, *Dn, *rows;
    double *w, *trs, *fx, *gx, *gx0, *min_cv, *gxc;
    uint8_t *upd, *live, *go;
    uint8_t *ph, *mu, *qdef;                 // ph: 0 = takes part in the round, 1 = QP running in the background, 2 = its QP answer arrived (joins at the step);
                                      // qdef: the round's capped QP launch handed the problem to the background launch;
                                      // mu: this iteration of the problem is one that may refresh the oracle at x (sip.py:855-870)
    int32_t *pit, *qdone;             // the problem's own outer-iteration index (a problem skips rounds while its QP is out); QP-answer flag
    int32_t *status, *niter, *ninst0;
    double *dxdes, *dx, *Rn, *tn, *fxn, *gxn, *sorig, *metric;
    double *qW, *qlo, *qhi; int32_t* qk; int32_t* qstatus;
    int32_t* list; unsigned long long* counters;      // [0] sweep list length, [1] live problems, [2] outer iterations, [3] slot overflows,
                                                      // [4] problems doing anything this round, [5] problems stopped because their QP had no solution
    double *T_rel, *bound, *dmin; int64_t* arg;
};

void ls_launch_init(const LsState& S, const LsSettings& c, cudaStream_t st);
void ls_launch_begin(const LsState& S, const LsSettings& c, cudaStream_t st);
void ls_launch_values(const LsState& S, int next, int want_rows, int write_d, cudaStream_t st);
void ls_launch_oracle_pre(const LsState& S, const LsSettings& c, cudaStream_t st);
void ls_launch_oracle_post(const LsState& S, const LsSettings& c, int64_t n, cudaStream_t st);
void ls_launch_gx(const LsState& S, const LsSettings& c, cudaStream_t st);
void ls_launch_step(const LsState& S, const LsSettings& c, cudaStream_t st);
void ls_launch_score_pre(const LsState& S, cudaStream_t st);
void ls_launch_score_post(const LsState& S, const LsSettings& c, int64_t n, cudaStream_t st);
void ls_launch_unpack_poses(const LsState& S, const double* raw /* [2][B][12] device: initial, desired */, cudaStream_t st);
void ls_launch_pack_params(const LsState& S, const int64_t* offs /* [B] device */, double* packed, cudaStream_t st);

}  // namespace sio
