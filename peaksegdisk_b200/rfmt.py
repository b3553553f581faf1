"""R's number formatting, needed bit for bit because the reference names its output files with
paste(penalty) (R/PeakSegFPOP_dir.R:64, R/sequentialSearch_dir.R:89-98).  Pure Python, no dependency
on the CUDA library: bench.py's reference arm and the golden generators import this module alone."""
import math

import numpy as np


def r_paste(x):
    """R's paste()/as.character() of one number: up to 15 significant digits, fixed notation unless
    scientific is narrower (R's formatReal rule), 'Inf' for infinity.  The sequential search names
    its output files with these strings, so they must match R's exactly."""
    if isinstance(x, str):
        return x
    if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
        return str(int(x))
    x = float(x)
    if math.isnan(x):
        return "NA"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0:
        return "0"
    neg = x < 0
    mant, exp10 = ("%.14e" % abs(x)).split("e")
    e10 = int(exp10)
    digits = mant.replace(".", "").rstrip("0") or "0"
    nsig = len(digits)
    # scientific width
    w_sci = (nsig + 1 if nsig > 1 else 1) + (4 if abs(e10) < 100 else 5)
    # fixed width
    left = e10 + 1 if e10 >= 0 else 1
    rgt = max(0, nsig - e10 - 1)
    w_fix = left + (rgt + 1 if rgt else 0)
    if w_fix <= w_sci:
        s = "%.*f" % (rgt, abs(x))
    else:
        s = "%.*e" % (nsig - 1, abs(x))
        m, e = s.split("e")
        s = "%se%s%02d" % (m, "-" if int(e) < 0 else "+", abs(int(e)))
    return ("-" if neg else "") + s
