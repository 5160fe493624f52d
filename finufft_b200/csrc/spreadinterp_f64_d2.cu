// Instantiates the spread / interp kernels for double, 2D, every supported kernel width.
#include "spreadinterp.cuh"
#define B200_NS_LIST B200_NS_CASE(2) B200_NS_CASE(3) B200_NS_CASE(4) B200_NS_CASE(5) B200_NS_CASE(6) B200_NS_CASE(7) B200_NS_CASE(8) B200_NS_CASE(9) B200_NS_CASE(10) B200_NS_CASE(11) B200_NS_CASE(12) B200_NS_CASE(13) B200_NS_CASE(14) B200_NS_CASE(15) B200_NS_CASE(16)
namespace b200 {
B200_DEFINE_LAUNCH(double, 2)
}
