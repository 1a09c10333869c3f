# oracle/ref_build.mk -- TEST INFRASTRUCTURE ONLY.
#
# Builds oracle/_ref/liblpopc_ref.so: the reference's OWN transcription sources, compiled
# unmodified from where they lie under $(REF) (= /root/reference/Lpopc/src) against the
# Armadillo/IPOPT stand-in headers of oracle/ref_shim/, plus oracle/ref_driver.cpp.
# The reference's own build (CMake + Armadillo + IPOPT + MUMPS + OpenBLAS under ../ThirdParty,
# none of which is in the reference tree or this image) is NOT run.  No reference source is
# copied into this repository; outputs go to oracle/_ref/ only (git-ignored, travels with gpurun).
#
#   make -f oracle/ref_build.mk            (from the repo root; needs /root/reference)
REF ?= /root/reference/Lpopc/src
OUT := oracle/_ref
CXX ?= g++
# -include functional: the vendored spdlog (Core/spdlog) predates compilers that stopped
# including <functional> transitively.  -ffp-contract=off as for the restatement and the device.
CXXFLAGS := -std=c++14 -O2 -w -fPIC -ffp-contract=off -include functional \
            -Ioracle/ref_shim -I$(REF)/Common -I$(REF)/Core -I$(REF)/SparseMatrix
REF_SRCS := SparseMatrix/LpSparseMatrix.cpp SparseMatrix/LpSparseArray.cpp \
            Core/RPMGenerator.cpp Core/LpSizeChecker.cpp Core/LpBoundsChecker.cpp Core/LpOptimalProblem.cpp \
            Core/LpGuessChecker.cpp Core/LpDerivDependciesChecker.cpp Core/LpFiniteDifferenceDerive.cpp \
            Core/LpNLPWrapper.cpp Core/LpHessian.cpp Core/LpSacleOCP.cpp Core/LpSolutionError.cpp Core/LpPhMeshRefineAlg.cpp Core/LpLiuHpMeshRefineAlg.cpp Core/Nlp2OPConverter.cpp \
            Common/LpOption.cpp Common/LpOptionList.cpp Common/LpReporter.cpp Common/LpDebug.cpp Common/LpUtils.cpp
# the reference's example programs: their user-function classes are the second opinion on include/problems/*.h.
# -Dmain=...: each example's main() becomes an unused internal function (it needs IPOPT to link)
EXAMPLES ?= /root/reference/Lpopc/example
EX_SRCS := hypersensitive/HyperSensitive.cpp bryson-denham/BrysonDenham.cpp launch/Launch.cpp
EX_INC := -I$(EXAMPLES)/hypersensitive -I$(EXAMPLES)/bryson-denham -I$(EXAMPLES)/launch
OBJS := $(addprefix $(OUT)/,$(notdir $(REF_SRCS:.cpp=.o))) $(addprefix $(OUT)/ex_,$(notdir $(EX_SRCS:.cpp=.o))) \
        $(OUT)/ref_driver.o $(OUT)/ref_examples.o
HDRS := oracle/ref_shim/armadillo $(wildcard include/*.h) $(wildcard include/problems/*.h)
vpath %.cpp $(REF)/Core $(REF)/Common $(REF)/SparseMatrix oracle $(EXAMPLES)/hypersensitive $(EXAMPLES)/bryson-denham $(EXAMPLES)/launch

$(OUT)/liblpopc_ref.so: $(OBJS)
	$(CXX) -shared -pthread -o $@ $(OBJS)

$(OUT)/ex_%.o: %.cpp $(HDRS) | $(OUT)
	$(CXX) $(CXXFLAGS) -Dmain="static lpopc_example_main" -c $< -o $@

$(OUT)/ref_examples.o: ref_examples.cpp $(HDRS) | $(OUT)
	$(CXX) $(CXXFLAGS) $(EX_INC) -c $< -o $@

$(OUT)/%.o: %.cpp $(HDRS) | $(OUT)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(OUT):
	mkdir -p $(OUT)

clean:
	rm -rf $(OUT)
.PHONY: clean
