# Build of the B200 MPPI core (product), the CPU oracle (test infrastructure) and, when /root/reference is
# mounted, the unmodified reference translation units against stub ROS/tf/Eigen headers (oracle/_ref).
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := /usr/bin/g++
CC        := /usr/bin/gcc
PKG       := ccv_mppi_path_tracker_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/libmppi_b200.so
ARCH      := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: every FMA on the device is an explicit fmaf() (FP32 contract of mppi_math.h)
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas -Xptxas -v
HOSTFP    := -ffp-contract=off -mfma
REF       ?= /root/reference

all: lib oracle harness hosttest

lib: $(LIB)
$(LIB): $(CSRC)/mppi_kernels.cu $(CSRC)/mppi_rollout_pruned.cu $(CSRC)/mppi_device.cuh $(CSRC)/mppi_capi.cu $(CSRC)/mppi_kernels.h $(CSRC)/mppi_math.h $(CSRC)/mppi_host.h $(CSRC)/philox.h include/mppi_b200.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/mppi_kernels.cu $(CSRC)/mppi_rollout_pruned.cu $(CSRC)/mppi_capi.cu -ldl 2> $(CSRC)/ptxas.log || (cat $(CSRC)/ptxas.log; exit 1)
	@grep -E "registers|spill" $(CSRC)/ptxas.log | sort | uniq -c | sort -rn | head -40 || true

harness: $(PKG)/mppi_harness
$(PKG)/mppi_harness: $(CSRC)/host/mppi_harness.cpp $(CSRC)/host/controllers.hpp $(LIB)
	$(CXX) -O2 -std=c++17 -Wall -o $@ $(CSRC)/host/mppi_harness.cpp -L$(PKG) -lmppi_b200 -Wl,-rpath,'$$ORIGIN'

hosttest: tests/host/fb_monitor_check tests/host/cmd_check
tests/host/cmd_check: tests/host/cmd_check.cpp $(CSRC)/host/controllers.hpp include/mppi_b200.h $(LIB)
	$(CXX) -O1 -std=c++17 -ffp-contract=off -Wall -o $@ tests/host/cmd_check.cpp -L$(PKG) -lmppi_b200 -Wl,-rpath,'$$ORIGIN/../../$(PKG)'
tests/host/fb_monitor_check: tests/host/fb_monitor_check.cpp $(CSRC)/host/controllers.hpp include/mppi_b200.h
	$(CXX) -O1 -std=c++17 -ffp-contract=off -Wall -o $@ tests/host/fb_monitor_check.cpp -L$(PKG) -lmppi_b200 -Wl,-rpath,'$$ORIGIN/../../$(PKG)'

oracle: oracle/liboracle.so oracle/libtwin.so
oracle/liboracle.so: oracle/mppi_oracle.c oracle/mppi_oracle.h
	$(CC) -O2 -std=gnu11 -fPIC -shared -fopenmp -ffp-contract=off -Wall -o $@ oracle/mppi_oracle.c -lm
oracle/libtwin.so: oracle/mppi_twin.cpp $(CSRC)/mppi_math.h $(CSRC)/mppi_host.h
	$(CXX) -O2 -std=c++17 -fPIC -shared $(HOSTFP) -Wall -Wno-unknown-pragmas -o $@ oracle/mppi_twin.cpp

# Unmodified reference TUs, compiled where they lie, against the stub headers in oracle/ref_shim.
ref:
	@if [ -d $(REF)/src ]; then $(MAKE) -C oracle/ref_shim REF=$(REF); else echo "reference not mounted: skipping oracle/_ref"; fi

clean:
	rm -f $(LIB) $(PKG)/mppi_harness oracle/*.so oracle/_ref/* $(CSRC)/ptxas.log

.PHONY: all lib oracle ref harness hosttest clean
