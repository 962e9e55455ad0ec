/* Same include path as the reference's include/misc3d/segmentation/iterative_plane_segmentation.h; the B200 build provides the
 * classes of this header through the C-ABI facade. */
#pragma once
#include "../b200_facade.hpp"
