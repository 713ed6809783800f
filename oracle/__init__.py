"""CPU oracle for the RSoptSC hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy/scipy FP64 restatement of the reference's R algorithms
(R/SimilarityM.R, R/Clustering.R), used only as the *checker*:

  * tests/                      (parity of the CUDA path against it)
  * __graft_entry__.smoke()     (one small parity check on cuda:0)
  * bench.py                    (`cpu_baseline` leg and `--impl reference`)

Nothing under rsoptsc_b200/ imports this package; the product path fails loudly
when the CUDA library is missing instead of falling back to it.

Pinning: see oracle/rsoptsc_oracle.py header.
"""
