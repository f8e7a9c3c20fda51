"""CPU-only checks of the host-side scheduling logic of the drop-in interface (no kernel is launched)."""
import pytest


def test_pass_plan_covers_every_stack_once():
    from upliftingtabletennis_b200.interface import _Detector
    for n in (0, 1, 3, 15, 16, 17, 32, 33, 100, 288):
        for streaming in (False, True):
            plan = _Detector._pass_plan(n, 16, streaming)
            assert sum(ns for _, ns in plan) == n
            pos = 0
            for s0, ns in plan:
                assert s0 == pos and 1 <= ns <= 16
                pos += ns
    # frames still arriving: two short passes first (the network starts after 6 frames), then full chunks
    assert _Detector._pass_plan(32, 16, True) == [(0, 4), (4, 12), (16, 16)]
    assert _Detector._pass_plan(16, 16, True) == [(0, 4), (4, 12)]
    assert _Detector._pass_plan(32, 16, False) == [(0, 16), (16, 16)]
    assert _Detector._pass_plan(10, 16, True) == [(0, 10)]          # short clips are one pass


def test_unsupported_detector_names_raise():
    """SegFormer++ sources are not part of the reference tree (SURVEY.md finding 2): a clear error, no silent substitute."""
    from upliftingtabletennis_b200 import interface
    with pytest.raises(NotImplementedError):
        raise interface._unsupported('segformerpp_b2')


def test_frame_key_follows_memory_not_objects():
    """A sliding window over a clip names every frame three times through fresh view objects: they must map to one upload."""
    import numpy as np
    import torch
    from upliftingtabletennis_b200.interface import _Detector
    clip = np.zeros((5, 4, 6, 3), np.uint8)
    assert _Detector._frame_key(clip[2]) == _Detector._frame_key(clip[2])
    assert clip[2] is not clip[2]
    assert len({_Detector._frame_key(clip[i]) for i in range(5)}) == 5
    assert _Detector._frame_key(clip[1]) != _Detector._frame_key(clip[1][:, :3])       # a crop is a different frame
    t = torch.from_numpy(clip)
    assert _Detector._frame_key(t[3]) == _Detector._frame_key(t[3])
    triples = [(t[i], t[i + 1], t[i + 2]) for i in range(3)]
    assert len({_Detector._frame_key(f) for tr in triples for f in tr}) == 5
