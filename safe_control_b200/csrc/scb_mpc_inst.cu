// scb_mpc_inst.cu -- one MPC-CBF model per translation unit: compiled once per model with -DSCB_MPC_INST=<scb_model id>
// (safe_control_b200/build.py), so the four heavy kernel instantiations build in parallel.
#define SCB_MPC_NO_DISPATCH
#include "scb_mpc_impl.cuh"

#ifndef SCB_MPC_INST
#error "compile with -DSCB_MPC_INST=<model id>"
#endif

namespace scb {
SCB_MPC_INSTANTIATE(SCB_MPC_INST)
}
