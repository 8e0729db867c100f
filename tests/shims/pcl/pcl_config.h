// TEST-ONLY shim of pcl/pcl_config.h. -DSHIM_PCL_MINOR=10 gives the PCL 1.10 flavour (boost::shared_ptr cloud pointers,
// DLIO's Ubuntu 20.04 image), the default is 1.12 (std::shared_ptr). Not shipped.
#pragma once
#ifndef SHIM_PCL_MINOR
#define SHIM_PCL_MINOR 12
#endif
#define PCL_MAJOR_VERSION 1
#define PCL_MINOR_VERSION SHIM_PCL_MINOR
#define PCL_REVISION_VERSION 0
#define PCL_VERSION_CALC(MAJ, MIN, PATCH) (MAJ * 100000 + MIN * 100 + PATCH)
#define PCL_VERSION PCL_VERSION_CALC(PCL_MAJOR_VERSION, PCL_MINOR_VERSION, PCL_REVISION_VERSION)
#define PCL_VERSION_COMPARE(OP, MAJ, MIN, PATCH) (PCL_VERSION OP PCL_VERSION_CALC(MAJ, MIN, PATCH))
