// step_lr91.cu -- instantiates the fused step kernel for one model
// ({2D,3D} x {iso,aniso} x {plain,tracker}); see step_kernel.cuh / models.cuh.
#include "step_kernel.cuh"

namespace fwb {
FWB_DEFINE_MODEL_ENTRY(g_entry_luo_rudy91, Model<FWB_MODEL_LUO_RUDY91>)
}
