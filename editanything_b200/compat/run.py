"""Run one of the reference's scripts UNCHANGED against this backend:

    python -m editanything_b200.compat.run /path/to/editany_nogradio.py [script args...]

The compatibility modules (this directory) go first on sys.path, then the script runs as __main__ through runpy -
for a plain script file runpy leaves sys.path alone, so `from editany_lora import EditAnythingLoraModel`
(editany_nogradio.py:2) resolves here instead of to the reference checkout's own module of that name."""
import os
import runpy
import sys

COMPAT_DIR = os.path.dirname(os.path.abspath(__file__))


def run_script(path, argv=None):
    """Returns the script's globals (e.g. `refined, output, ref, text` of editany_nogradio.py:15)."""
    path = os.path.abspath(path)
    old_path, old_argv = list(sys.path), list(sys.argv)
    stale = [m for m in ("editany_lora", "segment_anything", "utils", "annotator") if m in sys.modules]
    saved = {m: sys.modules.pop(m) for m in stale}
    try:
        sys.path.insert(0, COMPAT_DIR)
        sys.argv = [path] + list(argv or [])
        return runpy.run_path(path, run_name="__main__")
    finally:
        sys.path[:] = old_path
        sys.argv[:] = old_argv
        for m in ("editany_lora", "segment_anything", "utils", "annotator"):
            sys.modules.pop(m, None)
            for k in [k for k in sys.modules if k.startswith(m + ".")]:
                sys.modules.pop(k, None)
        sys.modules.update(saved)


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    run_script(sys.argv[1], sys.argv[2:])


if __name__ == "__main__":
    main()
