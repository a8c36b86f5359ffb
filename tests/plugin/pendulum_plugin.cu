// TEST FIXTURE -- a user problem compiled into ITS OWN shared library (include/nmpc_b200/plugin.h, steps 2-3).
#include <nmpc_b200/plugin.h>

#include "pendulum.h"

NMPC_B200_REGISTER_DDP_MODEL("pendulum", Pendulum<double>);
