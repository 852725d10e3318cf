"""Drop-in for the reference's ``rewiring/rewire.py`` (rewiring/rewire.py:7-14): curvature-type dispatch."""
from rewiring.sdrf_cuda_bfc import sdrf_cuda_bfc


def rewire(dt, curv_type, max_iterations, removal_bound, tau):
    if curv_type == 'bfc':
        dt = sdrf_cuda_bfc(dt, loops=max_iterations, remove_edges=True,
                           removal_bound=removal_bound, tau=tau, is_undirected=True)
    elif curv_type is not None:
        # classical curvatures ('1d', 'augmented', 'haantjes') stay on the reference's own CPU implementation
        # (rewiring/sdrf_no_cuda.py), found through the merged package path when the reference is on sys.path.
        try:
            from rewiring.sdrf_no_cuda import sdrf_no_cuda
        except ImportError as exc:
            raise NotImplementedError(
                f"curv_type={curv_type!r}: only 'bfc' runs on the B200 path; put the reference checkout on "
                "sys.path after this package to use its sdrf_no_cuda") from exc
        dt = sdrf_no_cuda(dt, curv_type, loops=max_iterations, remove_edges=True,
                          removal_bound=removal_bound, tau=tau)
    return dt.edge_index
