"""Extract the reference's `joostTest` fixture (R/sysdata.rda, MIT licence, (c) 2019
Matt Karikomi -- see tests/golden/LICENSE.reference) into a small .npz that travels
to the GPU box.  Run in the build container only:

    python tests/golden/make_golden.py            # needs /root/reference

Stored fields (only the ones SURVEY.md section 8c uses as parity pins):
  data_{indices,indptr,values,shape}  raw counts, genes x cells, CSC   (testL2R2.R:6)
  idx        n x K int32, 0-based: columns jj may use (D[jj, idx[jj]] == 0) (testL2R2.R:13)
  S_{rows,cols,vals}, S_n            symmetric similarity, COO, both triangles (testClusterNumber.R:5)
  H          720 x 9 NMF basis of S ; labels 1-based int32
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.rda_reader import read_rda, sparse_to_scipy, strip  # noqa: E402

REF = os.environ.get("RSOPTSC_REFERENCE", "/root/reference")


def main():
    jt = read_rda(os.path.join(REF, "R", "sysdata.rda"))["joostTest"]
    names = strip(jt["_attr"]["names"])
    f = dict(zip(names, strip(jt)))
    data = sparse_to_scipy(f["data"])
    D = sparse_to_scipy(f["D"]).toarray()
    S = sparse_to_scipy(f["S"]).tocoo()
    H = sparse_to_scipy(f["H"]).toarray()
    labels = np.asarray(strip(f["labels"]), dtype=np.int32)
    n = D.shape[0]
    allowed = [np.flatnonzero(D[j] == 0) for j in range(n)]
    K = len(allowed[0])
    assert all(len(a) == K for a in allowed)
    idx = np.stack(allowed).astype(np.int32)
    vals = data.data
    assert np.all(vals == np.round(vals)) and vals.max() < 2 ** 15
    out = os.path.join(HERE, "joost_fixture.npz")
    np.savez_compressed(
        out,
        data_indices=data.indices.astype(np.int32), data_indptr=data.indptr.astype(np.int32),
        data_values=vals.astype(np.int16), data_shape=np.asarray(data.shape, dtype=np.int64),
        idx=idx,
        S_rows=S.row.astype(np.int32), S_cols=S.col.astype(np.int32), S_vals=S.data, S_n=np.int64(S.shape[0]),
        H=H, labels=labels,
    )
    print("wrote", out, os.path.getsize(out), "bytes; data", data.shape, "nnz", data.nnz,
          "K", K, "S nnz", S.nnz, "H", H.shape)


if __name__ == "__main__":
    main()
