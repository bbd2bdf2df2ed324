"""Shim for `bytecode` (reference import: visualizer.py:1); get_local is inactive, names unused."""


class Bytecode:  # pragma: no cover
    @staticmethod
    def from_code(code):
        raise NotImplementedError


class Instr:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError
