"""Host mirror of FlatHeightMap<S>::updateHeightsByPointGenerationFunctor
(fclb_heightmap_build_host, reference geometry/heightmap/flat_heightmap-inl.h:249-272)
against the bottom layer of the reference's LayeredHeightMap built from the same points.
Host-only: no GPU, no compute call."""
import numpy as np
import pytest

import scenes


@pytest.mark.parametrize("half_shape,res", [(64, 0.01), (32, 0.025), (256, 0.004)])
def test_heights_match_reference(ref_oracle, half_shape, res):
    import fclb200 as fclb

    pts = scenes.terrain_points(6 * half_shape * half_shape, half_shape * res, seed=4200 + half_shape)
    hid = ref_oracle.heightmap_create(pts, res, half_shape)
    for dt, st in ((np.float32, fclb.F32), (np.float64, fclb.F64)):
        ref_h, upper = ref_oracle.heightmap_export(hid, dt, half_shape)
        ours = fclb.heightmap_build_host(pts, res, half_shape, st)
        assert ours.shape == ref_h.shape
        assert np.array_equal(ours, ref_h)
        assert upper == int(ours.max())
        assert (ours == 0).any() and (ours > 0).any()
