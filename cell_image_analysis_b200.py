"""Import shim: the package directory is ``cell-image-analysis_b200/`` (the
name the layout contract asks for); a hyphen is not importable, so this module
loads that directory's ``__init__.py`` under the importable name
``cell_image_analysis_b200`` and replaces itself in ``sys.modules``."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "cell-image-analysis_b200")
_spec = _u.spec_from_file_location(__name__, _os.path.join(_d, "__init__.py"),
                                   submodule_search_locations=[_d])
_m = _u.module_from_spec(_spec)
_sys.modules[__name__] = _m
_spec.loader.exec_module(_m)
