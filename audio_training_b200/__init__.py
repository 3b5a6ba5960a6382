"""Importable alias of the `audio-training_b200/` package directory (a hyphen cannot appear in an import name).
`import audio_training_b200` executes audio-training_b200/__init__.py as this package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "audio-training_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
