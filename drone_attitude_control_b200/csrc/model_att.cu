// Instantiates the solve kernels for the generated 3-D attitude model "att" (BNMPC_MODEL_ATT: position, velocity and attitude
// quaternion as states, total thrust and body rates as inputs - the north-star's nx = 10, nu = 4 OCP): one dense block,
// sensitivities per stage and SQP iteration.  No fused closed loop: its control step is bnmpc_step_for_x0.
#include "bnmpc_kernels.cuh"
BNMPC_DEFINE_MODEL_OPS_(bnmpc::Model_att, bnmpc::KIND_ATT, ops_att, false, false)
