"""Progress-output gating (reference ``src/cna/tools/_out.py:4-9``): ``print(..., file=out)`` goes
to stdout when ``show_progress`` is set and is discarded otherwise."""
import sys


class _Discard:
    def write(self, _):
        pass

    def flush(self):
        pass


_DISCARD = _Discard()


def select_output(allow):
    return sys.stdout if allow else _DISCARD
