"""Install rfdnet_b200 as the reference's `pointnet2_ops._ext`.

After `install()`, the UNMODIFIED reference files
external/pointnet2_ops_lib/pointnet2_ops/{pointnet2_utils,pointnet2_modules}.py resolve
`import pointnet2_ops._ext as _ext` (pointnet2_utils.py:8) to the sm_100a implementation.
"""
import sys
import types

from . import _ext


def install(reference_pkg_path=None):
    pkg = sys.modules.get('pointnet2_ops')
    if pkg is None:
        pkg = types.ModuleType('pointnet2_ops')
        pkg.__path__ = [reference_pkg_path] if reference_pkg_path else []
        sys.modules['pointnet2_ops'] = pkg
    pkg._ext = _ext
    sys.modules['pointnet2_ops._ext'] = _ext
    return _ext
