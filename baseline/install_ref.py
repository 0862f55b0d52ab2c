"""Installs the UNMODIFIED reference package into ``baseline/_ref/`` (git-ignored, travels with gpurun).

``python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
--target baseline/_ref /root/reference`` fails in this image: the project's build backend is
poetry-core, which is neither installed nor in the wheelhouse (outcome recorded in DESIGN.md §2).
The package is pure Python with no build step, so this script does what the wheel would do: it
copies ``eilev/**/*.py`` byte for byte from the read-only checkout into ``baseline/_ref/eilev/``.
``eilev.model.{v1,v2,utils}`` need only torch + transformers; ``eilev.data`` imports pytorchvideo,
which the image lacks, so a stub package with the three names ``eilev/data/utils.py`` imports is
written next to it (only used by the tests that execute the reference's import blocks).

Nothing under ``baseline/_ref`` is ever committed or imported by the product path; only
``bench.py --impl reference`` / the ``library_bar`` leg and the tests use it.
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
DEST = HERE / "_ref"

PYTORCHVIDEO_STUB = {
    "pytorchvideo/__init__.py": "",
    "pytorchvideo/data/__init__.py": (
        "class ClipSampler:\n"
        "    def __init__(self, clip_duration=0):\n"
        "        from fractions import Fraction\n"
        "        self._clip_duration = Fraction(clip_duration)\n"
        "        self._current_clip_index = 0\n"
        "        self._current_aug_index = 0\n"
        "    def reset(self):\n"
        "        pass\n"
        "class LabeledVideoDataset:\n"
        "    pass\n"),
    "pytorchvideo/data/clip_sampling.py": (
        "from typing import NamedTuple\n"
        "from fractions import Fraction\n"
        "class ClipInfo(NamedTuple):\n"
        "    clip_start_sec: Fraction\n"
        "    clip_end_sec: Fraction\n"
        "    clip_index: int\n"
        "    aug_index: int\n"
        "    is_last_clip: bool\n"),
    "pytorchvideo/data/video.py": "class VideoPathHandler:\n    pass\n",
}


def install(src: str | Path = "/root/reference", dest: Path = DEST) -> bool:
    src = Path(src)
    pkg = src / "eilev"
    if not pkg.is_dir():
        return False
    out = dest / "eilev"
    if out.exists():
        shutil.rmtree(out)
    for f in pkg.rglob("*.py"):
        tgt = out / f.relative_to(pkg)
        tgt.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(f, tgt)
    stub_root = dest / "_stubs"
    for rel, body in PYTORCHVIDEO_STUB.items():
        p = stub_root / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(body)
    (dest / "SOURCE").write_text(f"copied from {src} (yukw777/EILEV @ 41c461c), unmodified\n")
    return True


def available(dest: Path = DEST) -> bool:
    return (dest / "eilev" / "model" / "v2.py").is_file()


def add_to_path(dest: Path = DEST) -> None:
    """Makes ``import eilev`` resolve to the installed reference (and pytorchvideo to the stub when absent)."""
    if str(dest) not in sys.path:
        sys.path.insert(0, str(dest))
    try:
        import pytorchvideo  # noqa: F401
    except ImportError:
        sys.path.append(str(dest / "_stubs"))


if __name__ == "__main__":
    ok = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("installed" if ok else "reference checkout not found; nothing installed", DEST)
