/* nmpc_b200 -- DDP kernels for the centroidal-motion functor: n_x = 9, input dimension 16 or 0 along the horizon
   (TestDDPCentroidalMotion.cpp), padded to NU = 16. */
#include <nmpc_b200/models/centroidal_motion.h>

#include <nmpc_b200/engine/register.cuh>

NMPC_B200_REGISTER_DDP_MODEL("centroidal_motion", nmpc_b200::models::CentroidalMotion<double>);
