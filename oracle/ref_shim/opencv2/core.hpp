// OpenCV API stand-in for building the reference ORBextractor.cc as a checker; see ../cvshim.hpp (or ../../cvshim.hpp)
#include "../cvshim.hpp"
