/* TEST INFRASTRUCTURE (oracle/shim): stands in for AGG 2.4's <agg_basics.h>, which is not vendored by the
 * reference and not installed here. Everything lives in agg_shim.h. Drop the real AGG include directory in
 * front of this one on the include path (oracle/ref_build.sh AGG_INCLUDE=...) to build against the real thing. */
#include "agg_shim.h"
