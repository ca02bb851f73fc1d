// TEST INFRASTRUCTURE: what cmake's configure_file would generate from Thirdparty/g2o/config.h.in with G2O_USE_OPENMP OFF
// (the reference's default, Thirdparty/g2o/CMakeLists.txt:47) and a shared library build.  Found through
// -I refshim/g2o_cfg/a/b because g2o includes it as "../../config.h".
#ifndef G2O_CONFIG_H
#define G2O_CONFIG_H
#define G2O_SHARED_LIBS 1
#endif
