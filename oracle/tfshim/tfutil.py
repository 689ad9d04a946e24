"""ORACLE / test infrastructure: stand-in for the three `tfutil` symbols that the reference's `loss.py` touches
(/root/reference/loss.py:12), so that loss.py itself can be executed UNMODIFIED on oracle/tfshim.  The real
tfutil.py imports `imp` (gone in Python 3.12) and builds TF graph summaries; nothing of that is arithmetic.

  lerp          tfutil.py:41-43   a + (b - a) * t          (restated verbatim: it IS arithmetic of the loss)
  autosummary   tfutil.py:155-186 returns its value unchanged (TensorBoard side effect only)
"""


def lerp(a, b, t):
    return a + (b - a) * t


def autosummary(name, value):
    return value
