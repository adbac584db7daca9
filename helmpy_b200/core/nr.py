"""Newton-Raphson cross-check entry points (reference helmpy/core/nr.py:506,
nr_ds.py:547).  Out of scope as a GPU kernel: NR is only
the cross-check for HELM.  Implemented on the host with a sparse Jacobian."""


def nr(grid_data_file_path, Print_Details=False, Mismatch=1e-4, Scale=1, MaxIterations=15,
       Enforce_Qlimits=True, Results_FileName='', Save_results=False):
    raise NotImplementedError("nr cross-check: not built yet")


def nr_ds(grid_data_file_path, Print_Details=False, Mismatch=1e-4, Scale=1, MaxIterations=15,
          Enforce_Qlimits=True, DSB_model=True, Results_FileName='', Save_results=False):
    raise NotImplementedError("nr_ds cross-check: not built yet")
