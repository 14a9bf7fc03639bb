# Fragment for src/tce/ccsd_t/GNUmakefile, next to the existing `ifdef TCE_CUDA` block (:34-45).
# TCE_TRIPLES_B200=1 replaces sd_t_total.o / memory.o / hybrid.o (the reference's CUDA back-end) by libnwc_triples and
# adds the ISO_C_BINDING module + the native-tier driver.  The Fortran call sites of Tier 1 do not change.
ifdef TCE_TRIPLES_B200
  OBJ_OPTIMIZE += ccsd_t_gpu.o ccsd_t_singles_gpu.o ccsd_t_doubles_gpu.o nwc_triples_mod.o ccsd_t_b200.o cr_ccsd_t_b200.o
  LIB_DEFINES  += -DTCE_CUDA -DTCE_TRIPLES_B200
  NWC_TRIPLES_DIR ?= $(NWCHEM_TOP)/contrib/nwc_triples
  # sm_100a only: no fatbin for other architectures, no fallback path
  CUDA_ARCH    = -gencode arch=compute_100a,code=sm_100a
nwc_triples_mod.o: $(NWC_TRIPLES_DIR)/integration/nwc_triples_mod.F90
	$(FC) $(FOPTIONS) $(FOPTIMIZE) -c $< -o $@
ccsd_t_b200.o: $(NWC_TRIPLES_DIR)/integration/ccsd_t_b200.F90 nwc_triples_mod.o
	$(FC) $(FOPTIONS) $(FOPTIMIZE) $(LIB_INCLUDES) -c $< -o $@
cr_ccsd_t_b200.o: $(NWC_TRIPLES_DIR)/integration/cr_ccsd_t_b200.F90 nwc_triples_mod.o
	$(FC) $(FOPTIONS) $(FOPTIMIZE) $(LIB_INCLUDES) -c $< -o $@
$(NWC_TRIPLES_DIR)/nwchem_b200/lib/libnwc_triples.so:
	$(MAKE) -C $(NWC_TRIPLES_DIR)/nwchem_b200/csrc
endif

# Fragment for src/config/makefile.h, in the TCE_CUDA link block (:3744-3754).  The `.cu` pattern rule (:4145-4150)
# is not used: the library has its own Makefile (nwchem_b200/csrc/Makefile, nvcc -gencode arch=compute_100a,code=sm_100a).
ifdef TCE_TRIPLES_B200
  CORE_LIBS += -L$(NWC_TRIPLES_DIR)/nwchem_b200/lib -lnwc_triples -lcudart -lstdc++ -ldl -lgomp
endif
