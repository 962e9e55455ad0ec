/* Same include path as the reference's include/misc3d/common/ransac.h; the B200 build provides the
 * classes of this header through the C-ABI facade. */
#pragma once
#include "../b200_facade.hpp"
