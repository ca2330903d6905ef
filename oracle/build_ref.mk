# oracle/build_ref.mk -- TEST INFRASTRUCTURE, not product.
#
# Compiles the UNMODIFIED reference (samanseifi/Tahoe) sources where they lie
# under $(REF) (default /root/reference) with plain g++/gcc -- the reference's
# own CMake build system is NOT run.  Objects go to $(OBJ) (scratch), the
# only outputs kept are
#     oracle/_ref/tahoe            the reference executable (CPU baseline, kind="reference")
#     oracle/_ref/tahoe_dump       in-process dumper (full-precision RHS/LHS/fields; oracle/ref_dump.cpp)
#     oracle/_ref/compare          the reference's own benchmark comparator (benchmark_XML/comparator/src), used by the plugin tests
#     oracle/_ref/lib*.a           static libs the dumper / plugin demo link against
# oracle/_ref/ is git-ignored but travels to the GPU box.
#
# Flags restate what the reference's CMakeLists ask for (root CMakeLists.txt:
# -std=c++14 -fpermissive -O2, __EXPAT__, __SPOOLES__, __F2C__; tahoe/CMakeLists.txt:44
# __DEVELOPMENT__: ElementListT.cpp references development classes unconditionally,
# so libdevelopment is compiled too, with development/CMakeLists.txt's exclusions).
#
# usage: make -f oracle/build_ref.mk -j8
REF  ?= /root/reference
OBJ  ?= /tmp/tahoe_ref_obj
OUT  ?= $(abspath $(dir $(lastword $(MAKEFILE_LIST))))/_ref
CXX  := /usr/bin/g++
CC   := /usr/bin/gcc
OPT  ?= -O2

EXPAT_DIR := $(REF)/expat/expat/expat-1.95.7/lib
WARN := -w
DEFS := -D__EXPAT__ -D__SPOOLES__ -D__F2C__ -D__DEVELOPMENT__ -DNDEBUG

hdrdirs = $(sort $(dir $(shell find $(1) -name '*.h')))
INC_TOOLBOX := $(call hdrdirs,$(REF)/toolbox/src)
INC_TAHOE   := $(call hdrdirs,$(REF)/tahoe/src) $(REF)/tahoe/config/
INC_SPOOLES := $(call hdrdirs,$(REF)/spooles/src) $(REF)/spooles/inc/
INC_DEV     := $(call hdrdirs,$(REF)/development/src) $(REF)/development/config/
INC_F2C     := $(call hdrdirs,$(REF)/f2c/src) $(REF)/f2c/inc/
INCS := $(addprefix -I,$(INC_TOOLBOX) $(INC_TAHOE) $(INC_SPOOLES) $(INC_F2C) $(EXPAT_DIR) $(INC_DEV))

SRC_TOOLBOX := $(shell find $(REF)/toolbox/src -name '*.cpp')
SRC_TAHOE   := $(filter-out %_old.cpp %/main/main.cpp,$(shell find $(REF)/tahoe/src -name '*.cpp' -o -name '*.c'))
SRC_SPOOLES := $(shell find $(REF)/spooles/src -name '*.c' | grep -v -E '/[Tt]ests?/')
DEV_EXCL := /DEM_ellip3d/|/membrane_fluid_interaction/|_old\.cpp$$|_continuum\.cpp$$|/DEM_coupling/|/PMLElement/|/craig_enhanced_strain_loc/|/fiber_composite/|/meshfree_grad_plast/|/micromorphic/|/micromorphic2/|/micromorphic_curr_config/|/optimization/|/solid_fluid_mix/|/surface_CB[^/]*/
SRC_DEV     := $(shell find $(REF)/development/src -name '*.cpp' | grep -v -E '$(DEV_EXCL)')
SRC_F2C     := $(shell find $(REF)/f2c/src -name '*.c' | grep -v -E '/UNUSED/|/IO/')
SRC_EXPAT   := $(EXPAT_DIR)/xmlparse.c $(EXPAT_DIR)/xmlrole.c $(EXPAT_DIR)/xmltok.c

o = $(patsubst $(REF)/%,$(OBJ)/%.o,$(1))
OBJ_TOOLBOX := $(call o,$(SRC_TOOLBOX))
OBJ_TAHOE   := $(call o,$(SRC_TAHOE))
OBJ_SPOOLES := $(call o,$(SRC_SPOOLES))
OBJ_EXPAT   := $(call o,$(SRC_EXPAT))
OBJ_DEV     := $(call o,$(SRC_DEV))
OBJ_F2C     := $(call o,$(SRC_F2C))

LIBS := $(OUT)/libtahoe.a $(OUT)/libtoolbox.a $(OUT)/libdevelopment.a $(OUT)/libtahoe_spooles.a $(OUT)/libtahoe_expat.a $(OUT)/libtahoe_f2c.a
# libdevelopment is NOT whole-archived: the research tree holds same-named duplicates of a few
# libtahoe/toolbox translation units; only what ElementListT etc. reference is pulled in.
LINK := -Wl,--start-group -Wl,--whole-archive $(OUT)/libtahoe.a $(OUT)/libtoolbox.a -Wl,--no-whole-archive $(OUT)/libdevelopment.a \
        $(OUT)/libtahoe_spooles.a $(OUT)/libtahoe_expat.a $(OUT)/libtahoe_f2c.a -Wl,--end-group -fopenmp -lm -ldl

all: $(OUT)/tahoe $(OUT)/tahoe_dump $(OUT)/compare

$(OBJ)/incs.rsp:
	@mkdir -p $(OBJ)
	@echo $(INCS) > $@

$(OBJ)/%.cpp.o: $(REF)/%.cpp $(OBJ)/incs.rsp
	@mkdir -p $(dir $@)
	@$(CXX) -std=c++14 -fpermissive $(WARN) $(OPT) -fopenmp $(DEFS) @$(OBJ)/incs.rsp -c $< -o $@

$(OBJ)/expat/%.c.o: $(REF)/expat/%.c
	@mkdir -p $(dir $@)
	@$(CC) -std=gnu99 $(WARN) -O2 -DXML_NS -DXML_DTD -DHAVE_MEMMOVE -I$(EXPAT_DIR) -c $< -o $@

$(OBJ)/%.c.o: $(REF)/%.c $(OBJ)/incs.rsp
	@mkdir -p $(dir $@)
	@$(CC) -std=gnu99 $(WARN) -O2 $(DEFS) @$(OBJ)/incs.rsp -c $< -o $@

$(OUT)/libtoolbox.a: $(OBJ_TOOLBOX)
	@mkdir -p $(OUT); rm -f $@; echo $^ > $(OBJ)/toolbox.rsp; ar qcs $@ @$(OBJ)/toolbox.rsp
$(OUT)/libtahoe.a: $(OBJ_TAHOE)
	@mkdir -p $(OUT); rm -f $@; echo $^ > $(OBJ)/tahoe.rsp; ar qcs $@ @$(OBJ)/tahoe.rsp
$(OUT)/libtahoe_spooles.a: $(OBJ_SPOOLES)
	@mkdir -p $(OUT); rm -f $@; echo $^ > $(OBJ)/spooles.rsp; ar qcs $@ @$(OBJ)/spooles.rsp
$(OUT)/libdevelopment.a: $(OBJ_DEV)
	@mkdir -p $(OUT); rm -f $@; echo $^ > $(OBJ)/dev.rsp; ar qcs $@ @$(OBJ)/dev.rsp
$(OUT)/libtahoe_f2c.a: $(OBJ_F2C)
	@mkdir -p $(OUT); rm -f $@; echo $^ > $(OBJ)/f2c.rsp; ar qcs $@ @$(OBJ)/f2c.rsp
$(OUT)/libtahoe_expat.a: $(OBJ_EXPAT)
	@mkdir -p $(OUT); rm -f $@; ar qcs $@ $^

$(OUT)/tahoe: $(OBJ)/tahoe/src/main/main.cpp.o $(LIBS)
	$(CXX) -o $@ $< $(LINK)

REPO_ORACLE := $(abspath $(dir $(lastword $(MAKEFILE_LIST))))
$(OBJ)/ref_dump.o: $(REPO_ORACLE)/ref_dump.cpp $(OBJ)/incs.rsp
	$(CXX) -std=c++14 -fpermissive $(WARN) $(OPT) $(DEFS) @$(OBJ)/incs.rsp -c $< -o $@
$(OUT)/tahoe_dump: $(OBJ)/ref_dump.o $(LIBS)
	$(CXX) -o $@ $< $(LINK)

libs: $(LIBS)
.PHONY: all libs

# the reference's regression comparator (run_benchmarks.sh: `compare -f case.xml` checks case.io0.run against benchmark/case.io0.run
# with its default tolerances -- a value fails when it is off by more than 1e-8 relative AND 1e-10 absolute -- and prints "case.xml: PASS")
$(OUT)/compare: $(REF)/benchmark_XML/comparator/src/main.cpp $(REF)/benchmark_XML/comparator/src/ComparatorT.cpp $(OUT)/libtoolbox.a $(OUT)/libtahoe_expat.a $(OBJ)/incs.rsp
	@$(CXX) -std=c++14 -fpermissive $(WARN) $(OPT) -D__EXPAT__ -DNDEBUG @$(OBJ)/incs.rsp -I$(REF)/benchmark_XML/comparator/src \
	    $(REF)/benchmark_XML/comparator/src/main.cpp $(REF)/benchmark_XML/comparator/src/ComparatorT.cpp -o $@ \
	    -Wl,--start-group $(OUT)/libtoolbox.a $(OUT)/libtahoe_expat.a -Wl,--end-group -fopenmp -lm
