/* TEST INFRASTRUCTURE (oracle/shim): boost is not installed; boost::thread_interrupted lives in caffe/ofdg_caffe_shim.hpp */
#include "caffe/ofdg_caffe_shim.hpp"
