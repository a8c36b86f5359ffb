/* nmpc_b200 -- DDP kernels for the vertical-motion functor: n_x = 2, time-varying input dimension (0, 1 or 2 contact
   forces; TestDDPVerticalMotion.cpp), padded to NU = 2. */
#include <nmpc_b200/models/vertical_motion.h>

#include <nmpc_b200/engine/register.cuh>

NMPC_B200_REGISTER_DDP_MODEL("vertical_motion", nmpc_b200::models::VerticalMotion<double>);
