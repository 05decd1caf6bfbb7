"""
finitewave_b200 -- B200-native backend for finitewave's per-time-step hot path.

Drop-in for the on-path part of ``import finitewave as fw``: the same flat class
names (reference finitewave/__init__.py:18-109) for tissue, the six on-path
models, the isotropic / asymmetric stencils, stimuli, trackers and the loop's host
hooks.  The time loop itself (stimulus -> diffusion -> ionic -> trackers) runs as
hand-written sm_100a CUDA kernels behind the C ABI of include/finitewave_b200.h
(libfinitewave_b200.so, loaded with ctypes); PyTorch only owns device buffers.
There is no CPU fallback: without the CUDA library or a CUDA device every compute
entry raises ``FwbError``.
"""
import sys

from ._lib import FwbError
from .fibrosis import (Diffuse2DPattern, Diffuse3DPattern, FibrosisPattern, Structural2DPattern,
                       Structural3DPattern)
from .hooks import Command, CommandSequence, StateLoader, StateSaver, StateSaverCollection
from .model import (AlievPanfilov2D, AlievPanfilov3D, Barkley2D, Barkley3D, BuenoOrovio2D,
                    BuenoOrovio3D, CardiacModel, Courtemanche2D, Courtemanche3D,
                    FentonKarma2D, FentonKarma3D, LuoRudy912D, LuoRudy913D,
                    MitchellSchaeffer2D, MitchellSchaeffer3D, TP062D, TP063D)
from .stencil import (AsymmetricStencil2D, AsymmetricStencil3D, IsotropicStencil2D,
                      IsotropicStencil3D, Stencil, SymmetricStencil2D)
from .stimulation import (Stim, StimCurrent, StimCurrentArea2D, StimCurrentArea3D,
                          StimCurrentCoord2D, StimCurrentCoord3D, StimCurrentMatrix2D,
                          StimCurrentMatrix3D, StimSequence, StimVoltage, StimVoltageCoord2D,
                          StimVoltageCoord3D, StimVoltageListMatrix3D, StimVoltageMatrix2D,
                          StimVoltageMatrix3D)
from .tissue import (CardiacTissue, CardiacTissue2D, CardiacTissue3D, IncorrectNumberOfWeights,
                     IncorrectWeightsModeError2D, IncorrectWeightsShapeError)
from .tracker import (ActionPotential2DTracker, ActionPotential3DTracker, Animation2DTracker,
                      Animation3DTracker, AnimationSlice3DTracker,
                      ActivationTime2DTracker, ActivationTime3DTracker, ECG2DTracker,
                      ECG3DTracker, LocalActivationTime2DTracker, LocalActivationTime3DTracker,
                      MultiVariable2DTracker, MultiVariable3DTracker, Period2DTracker,
                      Period3DTracker, PeriodAnimation2DTracker, PeriodAnimation3DTracker,
                      SpiralWaveCore2DTracker, SpiralWaveCore3DTracker, Tracker,
                      TrackerSequence, Variable2DTracker, Variable3DTracker)

__version__ = "0.1.0"

# the reference's sub-module import paths (finitewave.cpuwave2D.fibrosis.diffuse_2d_pattern, ...)
from . import _compat  # noqa: E402
_compat.install(sys.modules[__name__])
