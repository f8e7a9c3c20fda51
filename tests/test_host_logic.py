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
