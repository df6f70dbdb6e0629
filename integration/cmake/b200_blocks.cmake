# cmake/Modules/b200_blocks.cmake -- included by the top-level CMakeLists.txt when ENABLE_B200=ON
# (integration/patches/gnss_block_factory_b200.patch).  Defines the `b200_blocks` target: the B200 blocks,
# their adapters and the host mirror of libb200gnss.so's C ABI.
#
#   cmake -DENABLE_B200=ON -DB200GNSS_ROOT=/path/to/b200-repo ..
#
# B200GNSS_ROOT is this repository: include/b200gnss.h, gnss_sdr_b200/libb200gnss.so (built with
# `python -m gnss_sdr_b200.build`), gnss_sdr_b200/host/*.cc and integration/src/**.
if(NOT B200GNSS_ROOT)
    message(FATAL_ERROR "ENABLE_B200=ON needs -DB200GNSS_ROOT=<path to the libb200gnss repository>")
endif()
find_library(B200GNSS_LIB b200gnss PATHS ${B200GNSS_ROOT}/gnss_sdr_b200 NO_DEFAULT_PATH REQUIRED)

set(B200_INTEGRATION ${B200GNSS_ROOT}/integration/src)
add_library(b200_blocks STATIC
    ${B200GNSS_ROOT}/gnss_sdr_b200/host/b200_multicorrelator_real_codes.cc
    ${B200GNSS_ROOT}/gnss_sdr_b200/host/b200_trk_coalescer.cc
    ${B200GNSS_ROOT}/gnss_sdr_b200/host/b200_pcps_acquisition_core.cc
    ${B200_INTEGRATION}/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking_b200.cc
    ${B200_INTEGRATION}/algorithms/tracking/adapters/b200_dll_pll_tracking.cc
    ${B200_INTEGRATION}/algorithms/acquisition/gnuradio_blocks/pcps_acquisition_b200.cc
    ${B200_INTEGRATION}/algorithms/acquisition/adapters/b200_pcps_acquisition.cc
)
target_include_directories(b200_blocks PUBLIC
    ${B200GNSS_ROOT}/include
    ${B200GNSS_ROOT}/gnss_sdr_b200/host
    ${B200_INTEGRATION}/algorithms/tracking/gnuradio_blocks
    ${B200_INTEGRATION}/algorithms/tracking/adapters
    ${B200_INTEGRATION}/algorithms/acquisition/gnuradio_blocks
    ${B200_INTEGRATION}/algorithms/acquisition/adapters
    ${B200_INTEGRATION}/core/receiver
)
# the same dependencies as the reference's own blocks of each kind
target_link_libraries(b200_blocks PUBLIC
    ${B200GNSS_LIB}
    tracking_libs
    acquisition_libs
    algorithms_libs
    channel_libs
    core_system_parameters
    Gnuradio::runtime
    Gnuradio::pmt
    Volkgnsssdr::volkgnsssdr
)
target_compile_definitions(b200_blocks PUBLIC -DB200_GPU_ACCEL=1)
target_compile_features(b200_blocks PUBLIC cxx_std_20)  # std::atomic wait/notify in the coalescer
