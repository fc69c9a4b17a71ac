"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithm for the exposure-synthesis path,
used as the parity checker.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the
product (wayne_b200/) never does.

  psf.py               ctypes access to the C restatement of PSF()
                       (psf_oracle.c -> _build/libpsf_oracle.so) and, when it was
                       built here, the UNMODIFIED reference kernel
                       (_ref/libwayne_ref_psf.so, _ref/pyparallel*.so)
  exposure_oracle.py   float64 numpy restatement of the reference's Python path
                       (exposure_generator.py / grism.py / detector.py / exposure.py)
"""
