# Merge with same-named packages later on sys.path (the reference checkout), so modules this package does not
# replace (e.g. curvature.classical_curvatures) keep resolving to the reference's files.
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
