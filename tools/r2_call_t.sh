#!/bin/bash
# round 2, call T: nn_distance candidate-group unroll / queries per thread
V=$PWD/monopsr_b200/build/variants
TAG=base timeout 60 python tools/nn_quick.py 2>&1 | tail -3
TAG=u4 MPB_LIB=$V/lib_nnu4.so timeout 60 python tools/nn_quick.py 2>&1 | tail -3
TAG=u8 MPB_LIB=$V/lib_nnu8.so timeout 60 python tools/nn_quick.py 2>&1 | tail -3
TAG=q1 MPB_NN_Q=1 timeout 60 python tools/nn_quick.py 2>&1 | tail -3
TAG=q1u4 MPB_NN_Q=1 MPB_LIB=$V/lib_nnu4.so timeout 60 python tools/nn_quick.py 2>&1 | tail -3
TAG=q4 MPB_NN_Q=4 timeout 60 python tools/nn_quick.py 2>&1 | tail -3
MPB_LIB=$V/lib_nnu4.so timeout 200 python -m pytest tests/test_tfops_gpu.py -q -x -p no:cacheprovider -k "nn_" 2>&1 | tail -1
