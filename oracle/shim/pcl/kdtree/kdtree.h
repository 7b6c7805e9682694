#ifndef ORACLE_SHIM_PCL_KDTREE_H
#define ORACLE_SHIM_PCL_KDTREE_H
#include <pcl/point_cloud.h>
#include <pcl/point_representation.h>
#endif
