#include "../refshim_cv.h"
#include "../refshim_filestorage.h"
