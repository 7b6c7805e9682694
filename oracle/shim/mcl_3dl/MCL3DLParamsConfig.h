// TEST INFRASTRUCTURE ONLY: stand-in for the catkin-generated dynamic_reconfigure config type.
#ifndef ORACLE_SHIM_MCL3DLPARAMSCONFIG_H
#define ORACLE_SHIM_MCL3DLPARAMSCONFIG_H
namespace mcl_3dl
{
struct MCL3DLParamsConfig
{
  double match_ratio_thresh, odom_err_integ_lin_sigma, odom_err_integ_ang_sigma;
};
}  // namespace mcl_3dl
#endif
