# same layout as the reference's generated package (python/CMakeLists.txt:27-29)
from .py_misc3d import *  # noqa: F401,F403
from .py_misc3d import common, registration, segmentation  # noqa: F401
