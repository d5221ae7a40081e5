#include "caffe/ofdg_caffe_shim.hpp"  /* TEST INFRASTRUCTURE (oracle/shim): Caffe is not vendored by the reference */
