"""TEST INFRASTRUCTURE.  A stand-in for the few python-stim classes the reference's own Python tests construct
(stim.DetectorErrorModel from text, stim.DemTarget, stim.DemInstruction, stim.target_logical_observable_id), so that
those test files can run unmodified against this repo's module on a box without stim
(tests/test_reference_python_tests.py).  Everything that would need stim's simulator (stim.Circuit) raises."""
from __future__ import annotations


def _norm(text: str) -> list[str]:
    return [" ".join(line.replace(",", ", ").split()) for line in str(text).splitlines() if line.strip()]


class DetectorErrorModel:
    def __init__(self, text: str = ""):
        self._lines = _norm(text)

    def __str__(self):
        return "\n".join(self._lines)

    def __repr__(self):
        return f"stim.DetectorErrorModel('''\n{self}\n''')"

    def __eq__(self, other):
        if isinstance(other, (DetectorErrorModel, str)):
            return _norm(str(self)) == _norm(str(other))
        return NotImplemented

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __hash__(self):
        return hash(tuple(self._lines))

    def _count(self, fn):
        return fn(str(self)) if self._lines else 0

    @property
    def num_detectors(self):
        from tesseract_decoder_b200 import capi
        return self._count(capi.dem_count_detectors)

    @property
    def num_observables(self):
        from tesseract_decoder_b200 import capi
        return self._count(capi.dem_count_observables)

    @property
    def num_errors(self):
        from tesseract_decoder_b200 import capi
        return self._count(capi.dem_count_errors)


class DemTarget:
    def __init__(self, text: str):
        self.text = str(text)

    def __str__(self):
        return self.text

    def __repr__(self):
        return f"stim.DemTarget('{self.text}')"

    def __eq__(self, other):
        return isinstance(other, DemTarget) and other.text == self.text

    def __hash__(self):
        return hash(self.text)


def target_logical_observable_id(index: int) -> DemTarget:
    return DemTarget(f"L{index}")


def target_relative_detector_id(index: int) -> DemTarget:
    return DemTarget(f"D{index}")


class DemInstruction:
    def __init__(self, type: str, args, targets, tag: str = ""):
        self.type, self.args, self.targets, self.tag = type, list(args), list(targets), tag

    def __str__(self):
        args = ", ".join(repr(float(a)) for a in self.args)
        return f"{self.type}({args}) " + " ".join(str(t) for t in self.targets)


class Circuit:
    def __init__(self, *a, **k):
        raise NotImplementedError("the stand-in has no circuit simulator; tests that need stim.Circuit are not run with it")

    @staticmethod
    def generated(*a, **k):
        raise NotImplementedError("the stand-in has no circuit simulator")


class CircuitInstruction:
    def __init__(self, *a, **k):
        raise NotImplementedError("the stand-in has no circuit simulator")
