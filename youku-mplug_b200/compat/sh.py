"""`sh.rm('-rf', path)` (utils.py:389) - the only use."""
import shutil


def rm(*args):
    for a in args:
        if not str(a).startswith("-"):
            shutil.rmtree(a, ignore_errors=True)
