#pragma once
/* stand-in for open3d/Macro.h (TEST INFRASTRUCTURE ONLY): the one macro misc3d/logging.h uses */
#define OPEN3D_FUNCTION __PRETTY_FUNCTION__
