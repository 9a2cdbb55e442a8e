# Builds the C-ABI library and its C / C++ clients without Python (what a Rust build.rs or a C++ project would run).
#   make            libimt_b200.so (nvcc, sm_100a only)
#   make clients    the plain-C client and the C++ mirror of the reference's tests (gcc / g++, no CUDA headers)
#   make oracle     the CPU oracle (test infrastructure)
# Python users: `python __graft_entry__.py` does the same through indexed-merkle-tree-halo2_b200/build.py.
NVCC    ?= /usr/local/cuda/bin/nvcc
PKG     := indexed-merkle-tree-halo2_b200
CSRC    := $(PKG)/csrc
OBJDIR  := $(PKG)/build
LIB     := $(PKG)/libimt_b200.so
UNITS   := imt_capi.cu imt_indexed.cu imt_spec.cu imt_latency.cu imt_comm.cu imt_io.cu poseidon_params.cpp
OBJS    := $(addprefix $(OBJDIR)/,$(addsuffix .o,$(basename $(UNITS))))
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 550 -Iinclude -I$(CSRC)
HEADERS := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h) include/imt_b200.h

all: $(LIB)

# carry discipline per kernel family (csrc/fr.cuh IMT_FREE_MASK; all 32 masks swept on a B200, profiles/r02_latency_lab.md)
$(OBJDIR)/imt_capi.o: NVFLAGS += -DIMT_FREE_MASK=22
$(OBJDIR)/imt_latency.o: NVFLAGS += -DIMT_FREE_MASK=29

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HEADERS)
	@mkdir -p $(OBJDIR)
	env -u CC -u CXX $(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJDIR)/%.o: $(CSRC)/%.cpp $(HEADERS)
	@mkdir -p $(OBJDIR)
	env -u CC -u CXX $(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	env -u CC -u CXX $(NVCC) -gencode arch=compute_100a,code=sm_100a -shared $(OBJS) -ldl -o $@

clients: $(LIB)
	@mkdir -p tests/_build
	gcc -std=c99 -Wall -Wextra -Werror -O1 -Iinclude tests/cabi_driver.c -o tests/_build/cabi_driver -L$(PKG) -limt_b200 -Wl,-rpath,$(abspath $(PKG))
	g++ -std=c++17 -Wall -Wextra -Werror -O1 -Iinclude tests/reference_tests.cpp -o tests/_build/reference_tests -L$(PKG) -limt_b200 -Wl,-rpath,$(abspath $(PKG))
	gcc -std=c99 -Wall -Wextra -Werror -O1 -Iinclude tests/cabi_multi_driver.c -o tests/_build/cabi_multi_driver -L$(PKG) -limt_b200 -Wl,-rpath,$(abspath $(PKG))

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(OBJDIR) $(LIB) tests/_build

.PHONY: all clients oracle clean
