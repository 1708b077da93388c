"""Importable alias of the package directory ``automatic-speech-recognition_b200``
(a hyphen cannot appear in an ``import`` statement):  ``import asr_b200``."""
import importlib
import sys

_pkg = importlib.import_module("automatic-speech-recognition_b200")
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_pkg.__name__ + "."):
        sys.modules[__name__ + _name[len(_pkg.__name__):]] = _mod
sys.modules[__name__] = _pkg
