// step_tp06.cu -- instantiates the fused step kernel for one model
// ({2D,3D} x {iso,aniso} x {plain,tracker}); see step_kernel.cuh / models.cuh.
#include "step_kernel.cuh"

namespace fwb {
FWB_DEFINE_MODEL_ENTRY(g_entry_tp06, Model<FWB_MODEL_TP06>)
}
