// step_nomodel.cu -- diffusion-only instantiation of the step kernel
// (fwb_diffuse: the reference's diffusion_kernel_{2d,3d}_{iso,aniso}).
#include "step_kernel.cuh"

namespace fwb {
FWB_DEFINE_MODEL_ENTRY(g_entry_nomodel, NoModel)
}
