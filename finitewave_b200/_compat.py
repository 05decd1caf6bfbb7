"""
Sub-module import paths of the reference, served lazily.

Almost all reference code says ``import finitewave as fw``, but the package's classes are also
importable by their defining module -- the reference's own fibrosis tests do
``from finitewave.cpuwave2D.fibrosis.diffuse_2d_pattern import Diffuse2DPattern``.  For a
drop-in the same paths must resolve here (and under whatever name the package is imported as,
e.g. ``sys.modules["finitewave"] = finitewave_b200``).  Instead of a hundred one-line files,
a meta-path finder synthesises ``<package>.<reference sub-module path>`` on demand from the
table below: a module object holding the classes this package provides for that path.

``_LAYOUT`` maps the reference's module tree below ``finitewave/`` (v0.8.5; ``tools`` and the
pyvista-only ``vtk_frame_3d_tracker`` excluded) to (is_package, exported class names); a
package also exports everything defined below it, as the reference's ``__init__`` files do.
Generated from the reference tree by listing class definitions per file; it holds names only.
"""
import importlib.abc
import importlib.machinery
import sys

_LAYOUT = {
    'core': (True, ['CardiacModel', 'CardiacTissue', 'Command', 'CommandSequence',
        'FibrosisPattern', 'IncorrectNumberOfWeights', 'IncorrectWeightsShapeError',
        'StateLoader', 'StateSaver', 'StateSaverCollection', 'Stencil', 'Stim', 'StimCurrent',
        'StimSequence', 'StimVoltage', 'Tracker', 'TrackerSequence']),
    'core.command': (True, ['Command', 'CommandSequence']),
    'core.exception': (True, ['IncorrectNumberOfWeights', 'IncorrectWeightsShapeError']),
    'core.fibrosis': (True, ['FibrosisPattern']),
    'core.model': (True, ['CardiacModel']),
    'core.state': (True, ['StateLoader', 'StateSaver', 'StateSaverCollection']),
    'core.stencil': (True, ['Stencil']),
    'core.stimulation': (True, ['Stim', 'StimCurrent', 'StimSequence', 'StimVoltage']),
    'core.tissue': (True, ['CardiacTissue']),
    'core.tracker': (True, ['Tracker', 'TrackerSequence']),
    'cpuwave2D': (True, ['ActionPotential2DTracker', 'ActivationTime2DTracker', 'AlievPanfilov2D',
        'Animation2DTracker', 'AsymmetricStencil2D', 'Barkley2D', 'BuenoOrovio2D',
        'CardiacTissue2D', 'Courtemanche2D', 'Diffuse2DPattern', 'ECG2DTracker', 'FentonKarma2D',
        'IncorrectWeightsModeError2D', 'IsotropicStencil2D', 'LocalActivationTime2DTracker',
        'LuoRudy912D', 'MitchellSchaeffer2D', 'MultiVariable2DTracker', 'Period2DTracker',
        'PeriodAnimation2DTracker', 'SpiralWaveCore2DTracker', 'StimCurrentArea2D',
        'StimCurrentCoord2D', 'StimCurrentMatrix2D', 'StimVoltageCoord2D', 'StimVoltageMatrix2D',
        'Structural2DPattern', 'SymmetricStencil2D', 'TP062D', 'Variable2DTracker']),
    'cpuwave2D.exception': (True, ['IncorrectWeightsModeError2D']),
    'cpuwave2D.fibrosis': (True, ['Diffuse2DPattern', 'Structural2DPattern']),
    'cpuwave2D.model': (True, ['AlievPanfilov2D', 'Barkley2D', 'BuenoOrovio2D', 'Courtemanche2D',
        'FentonKarma2D', 'LuoRudy912D', 'MitchellSchaeffer2D', 'TP062D']),
    'cpuwave2D.stencil': (True, ['AsymmetricStencil2D', 'IsotropicStencil2D',
        'SymmetricStencil2D']),
    'cpuwave2D.stimulation': (True, ['StimCurrentArea2D', 'StimCurrentCoord2D',
        'StimCurrentMatrix2D', 'StimVoltageCoord2D', 'StimVoltageMatrix2D']),
    'cpuwave2D.tissue': (True, ['CardiacTissue2D']),
    'cpuwave2D.tracker': (True, ['ActionPotential2DTracker', 'ActivationTime2DTracker',
        'Animation2DTracker', 'ECG2DTracker', 'LocalActivationTime2DTracker',
        'MultiVariable2DTracker', 'Period2DTracker', 'PeriodAnimation2DTracker',
        'SpiralWaveCore2DTracker', 'Variable2DTracker']),
    'cpuwave3D': (True, ['ActionPotential3DTracker', 'ActivationTime3DTracker', 'AlievPanfilov3D',
        'Animation3DTracker', 'AnimationSlice3DTracker', 'AsymmetricStencil3D', 'Barkley3D',
        'BuenoOrovio3D', 'CardiacTissue3D', 'Courtemanche3D', 'Diffuse3DPattern', 'ECG3DTracker',
        'FentonKarma3D', 'IsotropicStencil3D', 'LocalActivationTime3DTracker', 'LuoRudy913D',
        'MitchellSchaeffer3D', 'MultiVariable3DTracker', 'Period3DTracker',
        'PeriodAnimation3DTracker', 'SpiralWaveCore3DTracker', 'StimCurrentArea3D',
        'StimCurrentCoord3D', 'StimCurrentMatrix3D', 'StimVoltageCoord3D',
        'StimVoltageListMatrix3D', 'StimVoltageMatrix3D', 'Structural3DPattern', 'TP063D',
        'Variable3DTracker']),
    'cpuwave3D.fibrosis': (True, ['Diffuse3DPattern', 'Structural3DPattern']),
    'cpuwave3D.model': (True, ['AlievPanfilov3D', 'Barkley3D', 'BuenoOrovio3D', 'Courtemanche3D',
        'FentonKarma3D', 'LuoRudy913D', 'MitchellSchaeffer3D', 'TP063D']),
    'cpuwave3D.stencil': (True, ['AsymmetricStencil3D', 'IsotropicStencil3D']),
    'cpuwave3D.stimulation': (True, ['StimCurrentArea3D', 'StimCurrentCoord3D',
        'StimCurrentMatrix3D', 'StimVoltageCoord3D', 'StimVoltageListMatrix3D',
        'StimVoltageMatrix3D']),
    'cpuwave3D.tissue': (True, ['CardiacTissue3D']),
    'cpuwave3D.tracker': (True, ['ActionPotential3DTracker', 'ActivationTime3DTracker',
        'Animation3DTracker', 'AnimationSlice3DTracker', 'ECG3DTracker',
        'LocalActivationTime3DTracker', 'MultiVariable3DTracker', 'Period3DTracker',
        'PeriodAnimation3DTracker', 'SpiralWaveCore3DTracker', 'Variable3DTracker']),
    'core.command.command': (False, ['Command']),
    'core.command.command_sequence': (False, ['CommandSequence']),
    'core.exception.exceptions': (False, ['IncorrectNumberOfWeights',
        'IncorrectWeightsShapeError']),
    'core.fibrosis.fibrosis_pattern': (False, ['FibrosisPattern']),
    'core.model.cardiac_model': (False, ['CardiacModel']),
    'core.state.state_loader': (False, ['StateLoader']),
    'core.state.state_saver': (False, ['StateSaver', 'StateSaverCollection']),
    'core.stencil.stencil': (False, ['Stencil']),
    'core.stimulation.stim': (False, ['Stim']),
    'core.stimulation.stim_current': (False, ['StimCurrent']),
    'core.stimulation.stim_sequence': (False, ['StimSequence']),
    'core.stimulation.stim_voltage': (False, ['StimVoltage']),
    'core.tissue.cardiac_tissue': (False, ['CardiacTissue']),
    'core.tracker.tracker': (False, ['Tracker']),
    'core.tracker.tracker_sequence': (False, ['TrackerSequence']),
    'cpuwave2D.exception.exceptions_2d': (False, ['IncorrectWeightsModeError2D']),
    'cpuwave2D.fibrosis.diffuse_2d_pattern': (False, ['Diffuse2DPattern']),
    'cpuwave2D.fibrosis.structural_2d_pattern': (False, ['Structural2DPattern']),
    'cpuwave2D.model.aliev_panfilov_2d': (False, ['AlievPanfilov2D']),
    'cpuwave2D.model.barkley_2d': (False, ['Barkley2D']),
    'cpuwave2D.model.bueno_orovio_2d': (False, ['BuenoOrovio2D']),
    'cpuwave2D.model.courtemanche_2d': (False, ['Courtemanche2D']),
    'cpuwave2D.model.fenton_karma_2d': (False, ['FentonKarma2D']),
    'cpuwave2D.model.luo_rudy91_2d': (False, ['LuoRudy912D']),
    'cpuwave2D.model.mitchell_schaeffer_2d': (False, ['MitchellSchaeffer2D']),
    'cpuwave2D.model.tp06_2d': (False, ['TP062D']),
    'cpuwave2D.stencil.asymmetric_stencil_2d': (False, ['AsymmetricStencil2D']),
    'cpuwave2D.stencil.isotropic_stencil_2d': (False, ['IsotropicStencil2D']),
    'cpuwave2D.stencil.symmetric_stencil_2d': (False, ['SymmetricStencil2D']),
    'cpuwave2D.stimulation.stim_current_area_2d': (False, ['StimCurrentArea2D']),
    'cpuwave2D.stimulation.stim_current_coord_2d': (False, ['StimCurrentCoord2D']),
    'cpuwave2D.stimulation.stim_current_matrix_2d': (False, ['StimCurrentMatrix2D']),
    'cpuwave2D.stimulation.stim_voltage_coord_2d': (False, ['StimVoltageCoord2D']),
    'cpuwave2D.stimulation.stim_voltage_matrix_2d': (False, ['StimVoltageMatrix2D']),
    'cpuwave2D.tissue.cardiac_tissue_2d': (False, ['CardiacTissue2D']),
    'cpuwave2D.tracker.action_potential_2d_tracker': (False, ['ActionPotential2DTracker']),
    'cpuwave2D.tracker.activation_time_2d_tracker': (False, ['ActivationTime2DTracker']),
    'cpuwave2D.tracker.animation_2d_tracker': (False, ['Animation2DTracker']),
    'cpuwave2D.tracker.ecg_2d_tracker': (False, ['ECG2DTracker']),
    'cpuwave2D.tracker.local_activation_time_2d_tracker': (False,
        ['LocalActivationTime2DTracker']),
    'cpuwave2D.tracker.multi_variable_2d_tracker': (False, ['MultiVariable2DTracker']),
    'cpuwave2D.tracker.period_2d_tracker': (False, ['Period2DTracker']),
    'cpuwave2D.tracker.period_animation_2d_tracker': (False, ['PeriodAnimation2DTracker']),
    'cpuwave2D.tracker.spiral_wave_core_2d_tracker': (False, ['SpiralWaveCore2DTracker']),
    'cpuwave2D.tracker.variable_2d_tracker': (False, ['Variable2DTracker']),
    'cpuwave3D.fibrosis.diffuse_3d_pattern': (False, ['Diffuse3DPattern']),
    'cpuwave3D.fibrosis.structural_3d_pattern': (False, ['Structural3DPattern']),
    'cpuwave3D.model.aliev_panfilov_3d': (False, ['AlievPanfilov3D']),
    'cpuwave3D.model.barkley_3d': (False, ['Barkley3D']),
    'cpuwave3D.model.bueno_orovio_3d': (False, ['BuenoOrovio3D']),
    'cpuwave3D.model.courtemanche_3d': (False, ['Courtemanche3D']),
    'cpuwave3D.model.fenton_karma_3d': (False, ['FentonKarma3D']),
    'cpuwave3D.model.luo_rudy91_3d': (False, ['LuoRudy913D']),
    'cpuwave3D.model.mitchell_schaeffer_3d': (False, ['MitchellSchaeffer3D']),
    'cpuwave3D.model.tp06_3d': (False, ['TP063D']),
    'cpuwave3D.stencil.asymmetric_stencil_3d': (False, ['AsymmetricStencil3D']),
    'cpuwave3D.stencil.isotropic_stencil_3d': (False, ['IsotropicStencil3D']),
    'cpuwave3D.stimulation.stim_current_area_3d': (False, ['StimCurrentArea3D']),
    'cpuwave3D.stimulation.stim_current_coord_3d': (False, ['StimCurrentCoord3D']),
    'cpuwave3D.stimulation.stim_current_matrix_3d': (False, ['StimCurrentMatrix3D']),
    'cpuwave3D.stimulation.stim_voltage_coord_3d': (False, ['StimVoltageCoord3D']),
    'cpuwave3D.stimulation.stim_voltage_list_matrix_3d': (False, ['StimVoltageListMatrix3D']),
    'cpuwave3D.stimulation.stim_voltage_matrix_3d': (False, ['StimVoltageMatrix3D']),
    'cpuwave3D.tissue.cardiac_tissue_3d': (False, ['CardiacTissue3D']),
    'cpuwave3D.tracker.action_potential_3d_tracker': (False, ['ActionPotential3DTracker']),
    'cpuwave3D.tracker.activation_time_3d_tracker': (False, ['ActivationTime3DTracker']),
    'cpuwave3D.tracker.animation_3d_tracker': (False, ['Animation3DTracker']),
    'cpuwave3D.tracker.animation_slice_3d_tracker': (False, ['AnimationSlice3DTracker']),
    'cpuwave3D.tracker.ecg_3d_tracker': (False, ['ECG3DTracker']),
    'cpuwave3D.tracker.local_activation_time_3d_tracker': (False,
        ['LocalActivationTime3DTracker']),
    'cpuwave3D.tracker.multi_variable_3d_tracker': (False, ['MultiVariable3DTracker']),
    'cpuwave3D.tracker.period_3d_tracker': (False, ['Period3DTracker']),
    'cpuwave3D.tracker.period_animation_3d_tracker': (False, ['PeriodAnimation3DTracker']),
    'cpuwave3D.tracker.spiral_wave_core_3d_tracker': (False, ['SpiralWaveCore3DTracker']),
    'cpuwave3D.tracker.variable_3d_tracker': (False, ['Variable3DTracker']),
}


class _ReferencePathFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Consulted after the regular finders, i.e. only for names that are not real
    sub-modules of this package."""

    def __init__(self, package):
        self._package = package

    def find_spec(self, fullname, path=None, target=None):
        top, _, rest = fullname.partition(".")
        if rest not in _LAYOUT or sys.modules.get(top) is not self._package:
            return None
        return importlib.machinery.ModuleSpec(fullname, self, is_package=_LAYOUT[rest][0])

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        rest = module.__name__.partition(".")[2]
        is_package, names = _LAYOUT[rest]
        for name in names:
            setattr(module, name, getattr(self._package, name))
        module.__all__ = list(names)
        module.__doc__ = (f"Reference import path finitewave.{rest}, served by "
                          f"{self._package.__name__} (finitewave_b200/_compat.py).")
        if is_package:
            module.__path__ = []


def install(package):
    """Register the finder once (called at the end of finitewave_b200/__init__.py)."""
    for f in sys.meta_path:
        if isinstance(f, _ReferencePathFinder) and f._package is package:
            return f
    finder = _ReferencePathFinder(package)
    sys.meta_path.append(finder)
    return finder
