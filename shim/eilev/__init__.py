"""Drop-in ``eilev`` package: the hot-path modules resolve to the B200 implementation
(``eilev_b200``), everything else — ``eilev.data.frame``, ``eilev.data.ego4d``, ... — falls through
to the reference's own package when it is on ``sys.path`` behind this directory.

    PYTHONPATH=<repo>/shim:<repo>:<EILEV checkout> python scripts/general/train_v2.py ...

so ``scripts/general/train_v2.py:21-27`` and ``samples/eilev_generate_action_narration.py:10-12`` run
unchanged.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
