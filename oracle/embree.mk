# Compiles the reference's vendored Embree 3.6.1 (third-party/embree) -- BVH8/Triangle4 SAH build, packet traversal, the
# Moeller-Trumbore intersectors -- from its own sources WHERE THEY LIE into oracle/_ref/libgxy_embree_scene_ref.so, behind
# embree_scene_ref.cpp (ours: a plain C interface "commit a triangle scene, intersect a ray array with rtcIntersect8").
# This is a recipe we wrote (the source lists are the ones third-party/embree/kernels/CMakeLists.txt:35-189 and
# common/*/CMakeLists.txt name for a TASKING_INTERNAL, ISPC-less build with ISAs SSE2/SSE4.2/AVX/AVX2, which is what Galaxy's own
# build selects on an AVX2 host; flags as common/cmake/gnu.cmake:17-38,52-55); the reference's build system is not run, no reference
# source is copied, outputs go to oracle/_ref only.  Test infrastructure + the CPU arm of bench.py (kind "reference").
#   make -C oracle embree        (about 4 minutes on 8 cores)
# the system compiler with its shared libstdc++: the image's /opt/gcc wrapper links a second, static C++ runtime into the library,
# which crashes inside a Python process (threads + exceptions + iostreams), see the note in Makefile
CXX     := /usr/bin/g++
REF     ?= /root/reference
EMBREE  := $(REF)/third-party/embree
OBJ     := _ref/embree_obj
COMMONF := -O3 -DNDEBUG -std=c++11 -fPIC -fvisibility=hidden -fvisibility-inlines-hidden -fno-strict-aliasing -fno-tree-vectorize -w \
           -DTASKING_INTERNAL -DEMBREE_TARGET_SSE2 -DEMBREE_TARGET_SSE42 -DEMBREE_TARGET_AVX -DEMBREE_TARGET_AVX2 -DEMBREE_STATIC_LIB \
           -I$(EMBREE) -I$(EMBREE)/include -I_ref/gen/x -I_ref/gen
F_SSE2  := -msse2
F_SSE42 := -msse4.2
F_AVX   := -mavx
F_AVX2  := -mf16c -mavx2 -mfma -mlzcnt -mbmi -mbmi2

# --- common/ libraries (sys, math, simd, lexers, tasking, algorithms), lowest ISA
COMMON_SRCS := sys/sysinfo sys/alloc sys/filename sys/library sys/thread sys/string sys/regression sys/mutex sys/condition sys/barrier \
               math/constants simd/sse lexers/stringstream lexers/tokenstream tasking/taskschedulerinternal \
               algorithms/parallel_for algorithms/parallel_reduce algorithms/parallel_prefix_sum algorithms/parallel_for_for \
               algorithms/parallel_for_for_prefix_sum algorithms/parallel_partition algorithms/parallel_sort algorithms/parallel_set \
               algorithms/parallel_map algorithms/parallel_filter

# --- kernels/, lowest ISA (EMBREE_LOWEST_ISA), no subdivision surfaces, ray packets on
BASE_SRCS := common/device common/stat common/acceln common/accelset common/state common/rtcore common/rtcore_builder common/scene \
             common/alloc common/geometry common/scene_user_geometry common/scene_instance common/scene_triangle_mesh \
             common/scene_quad_mesh common/scene_curves common/scene_line_segments common/scene_grid_mesh common/scene_points \
             subdiv/bezier_curve subdiv/bspline_curve subdiv/catmullrom_curve \
             geometry/primitive4 geometry/instance_intersector geometry/curve_intersector_virtual builders/primrefgen \
             bvh/bvh bvh/bvh_statistics bvh/bvh4_factory bvh/bvh8_factory bvh/bvh_rotate bvh/bvh_refit bvh/bvh_builder \
             bvh/bvh_builder_hair bvh/bvh_builder_hair_mb bvh/bvh_builder_morton bvh/bvh_builder_sah bvh/bvh_builder_sah_spatial \
             bvh/bvh_builder_sah_mb bvh/bvh_builder_twolevel bvh/bvh_intersector1_bvh4 \
             bvh/bvh_intersector_hybrid4_bvh4 bvh/bvh_intersector_stream_bvh4 bvh/bvh_intersector_stream_filters

# --- per-ISA lists (the embree_files macro)
ISA_COMMON := geometry/instance_intersector geometry/curve_intersector_virtual bvh/bvh_intersector1_bvh4 \
              bvh/bvh_intersector_hybrid4_bvh4 bvh/bvh_intersector_stream_bvh4 bvh/bvh_intersector_stream_filters
ISA_BUILD  := common/scene_user_geometry common/scene_instance common/scene_triangle_mesh common/scene_quad_mesh common/scene_curves \
              common/scene_line_segments common/scene_grid_mesh common/scene_points bvh/bvh_refit bvh/bvh_builder bvh/bvh_builder_hair \
              bvh/bvh_builder_hair_mb bvh/bvh_builder_sah bvh/bvh_builder_sah_spatial bvh/bvh_builder_sah_mb bvh/bvh_builder_twolevel \
              bvh/bvh_builder_morton bvh/bvh_rotate builders/primrefgen
ISA_WIDE   := bvh/bvh_intersector1_bvh8 bvh/bvh_intersector_hybrid8_bvh4 bvh/bvh_intersector_hybrid4_bvh8 \
              bvh/bvh_intersector_hybrid8_bvh8 bvh/bvh_intersector_stream_bvh8
SSE42_SRCS := $(ISA_COMMON)
AVX_SRCS   := $(ISA_COMMON) geometry/primitive8 $(ISA_BUILD) $(ISA_WIDE) bvh/bvh bvh/bvh_statistics
AVX2_SRCS  := $(ISA_COMMON) $(ISA_BUILD) $(ISA_WIDE)

objs = $(addprefix $(OBJ)/$(1)/,$(addsuffix .o,$(2)))
ALL_OBJS := $(call objs,common,$(COMMON_SRCS)) $(call objs,base,$(BASE_SRCS)) $(call objs,sse42,$(SSE42_SRCS)) \
            $(call objs,avx,$(AVX_SRCS)) $(call objs,avx2,$(AVX2_SRCS))

embree: _ref/libgxy_embree_scene_ref.so

# kernels/common/device.cpp includes "../hash.h" (cmake writes it from hash.h.in with the git hash; a fixed string here)
_ref/gen/hash.h: $(EMBREE)/kernels/hash.h.in
	mkdir -p _ref/gen/x
	sed -e 's/@EMBREE_HASH@/vendored-3.6.1/' $< > $@
_ref/gen/config.h _ref/gen/rtcore_config.h:
	$(MAKE) -f Makefile $@

$(OBJ)/common/%.o: $(EMBREE)/common/%.cpp _ref/gen/config.h _ref/gen/rtcore_config.h _ref/gen/hash.h
	@mkdir -p $(dir $@)
	$(CXX) $(COMMONF) $(F_SSE2) -c $< -o $@
$(OBJ)/base/%.o: $(EMBREE)/kernels/%.cpp _ref/gen/config.h _ref/gen/rtcore_config.h _ref/gen/hash.h
	@mkdir -p $(dir $@)
	$(CXX) $(COMMONF) $(F_SSE2) -DEMBREE_LOWEST_ISA -c $< -o $@
$(OBJ)/sse42/%.o: $(EMBREE)/kernels/%.cpp _ref/gen/config.h _ref/gen/rtcore_config.h _ref/gen/hash.h
	@mkdir -p $(dir $@)
	$(CXX) $(COMMONF) $(F_SSE42) -c $< -o $@
$(OBJ)/avx/%.o: $(EMBREE)/kernels/%.cpp _ref/gen/config.h _ref/gen/rtcore_config.h _ref/gen/hash.h
	@mkdir -p $(dir $@)
	$(CXX) $(COMMONF) $(F_AVX) -c $< -o $@
$(OBJ)/avx2/%.o: $(EMBREE)/kernels/%.cpp _ref/gen/config.h _ref/gen/rtcore_config.h _ref/gen/hash.h
	@mkdir -p $(dir $@)
	$(CXX) $(COMMONF) $(F_AVX2) -c $< -o $@

_ref/libgxy_embree_scene_ref.so: embree_scene_ref.cpp $(ALL_OBJS)
	$(CXX) -O2 -std=c++14 -fPIC -pthread -DEMBREE_STATIC_LIB -w -I$(EMBREE)/include -I_ref/gen -shared -o $@.tmp embree_scene_ref.cpp $(ALL_OBJS) -lpthread -ldl \
	    && mv -f $@.tmp $@

.PHONY: embree
