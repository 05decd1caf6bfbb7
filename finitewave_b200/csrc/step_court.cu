// step_court.cu -- instantiates the fused step kernels for one model
// ({2D,3D} x {iso,aniso} x {plain,tracker} x {single,slab}); see step_kernel.cuh / models.cuh.
#include "step_kernel.cuh"

namespace fwb {
FWB_DEFINE_MODEL_ENTRY(g_entry_courtemanche, Model<FWB_MODEL_COURTEMANCHE>)
}
