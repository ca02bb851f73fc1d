#include "../../refshim_cv.h"
