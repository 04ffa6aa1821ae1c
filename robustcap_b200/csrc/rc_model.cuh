// rc_model: SMPL constants resident on the device (shared by kinematics.cu, fusion.cu, smplify.cu).
#pragma once
#include "rc_rows.h"

struct rc_model {
    RcModelConst host;              // host copy of the per-frame constants
    RcModelConst* d_const = nullptr;
    float* d_verts = nullptr;       // [nv,3] zero-pose vertices, root at the origin
    float* d_skin_w = nullptr;      // [nv,24]
    int nv = 0;
};
