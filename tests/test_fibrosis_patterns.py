"""
Fibrosis patterns: Diffuse{2,3}DPattern / Structural{2,3}DPattern of the
reference (finitewave/cpuwave2D/fibrosis/*.py, cpuwave3D/fibrosis/*.py) generated on the
device.  The reference draws from global RNG state, so parity is on the SEMANTICS (box
overwritten with 1 / 2, block anchoring and clipping, density) plus bit-exactness of the CUDA
kernel against the numpy restatement of the same hash draw (oracle.fibrosis_pattern).
"""
import numpy as np
import pytest

from oracle import oracle


def test_oracle_pattern_semantics():
    mesh = np.ones((40, 50), dtype=np.int8)
    mesh[:, :7] = 0                                           # empty strip: the box overwrites it
    out = oracle.fibrosis_pattern(mesh, [5, 33, 3, 47], [6, 8], 0.5, seed=3)
    assert set(np.unique(out[5:33, 3:47])) <= {1, 2}
    assert np.array_equal(out[:5], mesh[:5]) and np.array_equal(out[33:], mesh[33:])
    assert np.array_equal(out[:, :3], mesh[:, :3]) and np.array_equal(out[:, 47:], mesh[:, 47:])
    # blocks are anchored at the box origin, constant inside, clipped at the box end
    for i0 in range(5, 33, 6):
        for j0 in range(3, 47, 8):
            blk = out[i0:min(i0 + 6, 33), j0:min(j0 + 8, 47)]
            assert blk.min() == blk.max()
    # density (per block) and seed dependence
    big = oracle.fibrosis_pattern(np.ones((600, 600), np.int8), [0, 600, 0, 600], [1, 1], 0.3, 1)
    assert abs((big == 2).mean() - 0.3) < 0.005
    assert not np.array_equal(big, oracle.fibrosis_pattern(np.ones((600, 600), np.int8),
                                                           [0, 600, 0, 600], [1, 1], 0.3, 2))
    # empty box: untouched
    assert np.array_equal(oracle.fibrosis_pattern(mesh, [9, 9, 0, 50], [1, 1], 0.5), mesh)


@pytest.fixture(scope="module")
def fw():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import finitewave_b200
    return finitewave_b200


@pytest.mark.gpu
def test_device_patterns_match_oracle_bitwise(fw):
    m2 = np.ones((70, 90), dtype=np.int8)
    m2[10:20, :] = 0
    for pat, box, block in (
            (fw.Diffuse2DPattern(0.35, seed=5), [0, 70, 0, 90], [1, 1]),
            (fw.Diffuse2DPattern(0.2, 5, 60, 7, 81, seed=9), [5, 60, 7, 81], [1, 1]),
            (fw.Structural2DPattern(0.4, 4, 9, 3, 66, 2, 88, seed=1), [3, 66, 2, 88], [4, 9])):
        out = pat.generate(mesh=m2.copy())
        assert np.array_equal(out, oracle.fibrosis_pattern(m2, box, block, pat.density, pat.seed))
    m3 = np.ones((24, 30, 28), dtype=np.int8)
    for pat, box, block in (
            (fw.Diffuse3DPattern(2, 20, 0, 30, 5, 25, 0.25, seed=4), [2, 20, 0, 30, 5, 25], [1, 1, 1]),
            (fw.Structural3DPattern(1, 23, 2, 29, 0, 28, 0.3, 3, 5, 4, seed=8),
             [1, 23, 2, 29, 0, 28], [3, 5, 4])):
        out = pat.generate(mesh=m3.copy())
        assert np.array_equal(out, oracle.fibrosis_pattern(m3, box, block, pat.density, pat.seed))
    # shape= form and add_pattern (the tissue setter re-applies the empty outer ring)
    out = fw.Diffuse2DPattern(0.5, seed=2).generate(shape=(32, 32))
    assert out.dtype == np.int8 and set(np.unique(out)) == {1, 2}
    tissue = fw.CardiacTissue2D([64, 64])
    tissue.add_pattern(fw.Structural2DPattern(0.3, 5, 5, 10, 50, 10, 50, seed=6))
    assert (tissue.mesh[0] == 0).all() and (tissue.mesh[:, -1] == 0).all()
    assert (tissue.mesh[10:50, 10:50] == 2).any()


@pytest.mark.gpu
def test_device_pattern_is_partition_invariant(fw):
    """A slab generated with its slow_offset equals the same slices of the whole tissue."""
    import torch
    pat = fw.Structural3DPattern(3, 61, 0, 40, 4, 44, 0.35, 7, 3, 5, seed=11)
    full = torch.ones((64, 40, 48), dtype=torch.int8, device="cuda")
    pat.generate_device(full)
    for a, b in ((0, 17), (17, 40), (40, 64)):
        part = torch.ones((b - a, 40, 48), dtype=torch.int8, device="cuda")
        pat.generate_device(part, slow_offset=a, global_shape=(64, 40, 48))
        assert torch.equal(part, full[a:b])
    pat2 = fw.Diffuse2DPattern(0.3, seed=3)
    full2 = torch.ones((96, 64), dtype=torch.int8, device="cuda")
    pat2.generate_device(full2)
    part2 = torch.ones((40, 64), dtype=torch.int8, device="cuda")
    pat2.generate_device(part2, slow_offset=30, global_shape=(96, 64))
    assert torch.equal(part2, full2[30:70])
    assert abs(float((full2 == 2).float().mean()) - 0.3) < 0.03
