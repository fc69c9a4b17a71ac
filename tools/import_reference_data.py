"""Convert the reference's bundled instrument tables into this package's own
data files (run once, here, where /root/reference exists; outputs committed).

  wayne/data/wfc3_ir_mode_exptime.csv  -> wayne_b200/data/wfc3_ir_modes.json ["exptime"]
  wayne/data/wfc3_ir_mode_calb.csv     -> wayne_b200/data/wfc3_ir_modes.json ["dark"]
  wayne/data/wfc3_ir_initial_bias_256.fits (ext 1, 266x266 float64, integer
      valued) -> wayne_b200/data/wfc3_ir_initial_bias_256.npz (uint16, exact)

The CSVs are parsed exactly the way the reference does (wayne/detector.py:
248-267: pandas, skiprows=1, thousands=','), so the quirky rows
come out as the reference sees them.
"""
import json
import os
import sys

import numpy as np
import pandas as pd

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wayne_b200", "data")
sys.path.insert(0, os.path.dirname(OUT.rstrip("/")).rsplit("/", 1)[0])

exp = pd.read_csv(os.path.join(REF, "wayne/data/wfc3_ir_mode_exptime.csv"), skiprows=1,
                  dtype={"SUBARRAY": np.int64, "SAMPSEQ": object, "SAMPNUM": np.int64,
                         "TIME": float}, thousands=",")
cal = pd.read_csv(os.path.join(REF, "wayne/data/wfc3_ir_mode_calb.csv"), skiprows=1,
                  dtype={"SUBARRAY": np.int64, "SAMPSEQ": object, "dark": str})
doc = {
    "source": "HST Cycle 22 Phase II Proposal Instructions 13.3.6 / WFC3 ISR 2014-06, as "
              "tabulated by ucl-exoplanets/wayne (wayne/data/*.csv)",
    "exptime_columns": ["SUBARRAY", "SAMPSEQ", "SAMPNUM", "TIME"],
    "exptime": [[int(r.SUBARRAY), str(r.SAMPSEQ), int(r.SAMPNUM), float(r.TIME)]
                for r in exp.itertuples()],
    "dark_columns": ["SUBARRAY", "SAMPSEQ", "Dark"],
    "dark": [[int(r.SUBARRAY), str(r.SAMPSEQ), str(r.Dark)] for r in cal.itertuples()],
}
os.makedirs(OUT, exist_ok=True)
with open(os.path.join(OUT, "wfc3_ir_modes.json"), "w") as f:
    json.dump(doc, f, indent=0, separators=(",", ":"))

from wayne_b200 import fitsio  # noqa: E402

bias = fitsio.open(os.path.join(REF, "wayne/data/wfc3_ir_initial_bias_256.fits"))[1].data
b16 = bias.astype(np.uint16)
assert (b16.astype(np.float64) == bias).all()
np.savez_compressed(os.path.join(OUT, "wfc3_ir_initial_bias_256.npz"), bias=b16)
print(len(doc["exptime"]), len(doc["dark"]), bias.shape)
