// TEST INFRASTRUCTURE ONLY: parameters.h holds a unique_ptr<dynamic_reconfigure::Server<T>>.
#ifndef ORACLE_SHIM_DYNRECONF_H
#define ORACLE_SHIM_DYNRECONF_H
namespace dynamic_reconfigure
{
template <class T>
class Server
{
};
}  // namespace dynamic_reconfigure
#endif
