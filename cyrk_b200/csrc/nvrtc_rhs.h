// nvrtc_rhs.h -- user right-hand sides compiled at run time (declared here, defined in nvrtc_rhs.cpp).
#pragma once
