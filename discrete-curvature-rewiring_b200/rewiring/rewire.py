"""Drop-in for the reference's ``rewiring/rewire.py`` (rewiring/rewire.py:7-14): curvature-type dispatch.

``rewire(dt, curv_type, max_iterations, removal_bound, tau) -> edge_index``

* ``'bfc'``                              -> the device-resident BFC loop (``rewiring.sdrf_cuda_bfc``), undirected
* ``'1d'`` / ``'augmented'`` / ``'haantjes'`` -> the classical-curvature loop on the same device kernel
  (``rewiring.sdrf_no_cuda``)
* ``None``                               -> no rewiring: the input's ``edge_index`` comes back unchanged

Every flavour removes edges (``remove_edges=True``), as the reference's dispatcher does.
"""
from rewiring.sdrf_cuda_bfc import sdrf_cuda_bfc
from rewiring.sdrf_no_cuda import sdrf_no_cuda


def rewire(dt, curv_type, max_iterations, removal_bound, tau):
    if curv_type is None:
        return dt.edge_index
    common = dict(loops=max_iterations, remove_edges=True, removal_bound=removal_bound, tau=tau)
    if curv_type == 'bfc':
        rewired = sdrf_cuda_bfc(dt, is_undirected=True, **common)
    else:
        rewired = sdrf_no_cuda(dt, curv_type, **common)
    return rewired.edge_index
