# Goes to <sleqp>/cmake/SearchFactB200.cmake (pattern: cmake/SearchFactCHOLMOD.cmake:1-22).
# Once done, this will define
#
#  B200_INCLUDE_DIRS   - where to find sleqp_b200.h
#  B200_LIBRARIES      - libsleqp_b200.so (CUDA inside, C-ABI outside)
#  B200_VERSION        - version string of the backend
#  B200_FOUND          - True if the B200 backend was found.
#
# Hints: -DB200_DIR=<checkout of this repository> or the environment variable B200_DIR
# (headers in <dir>/include, library in <dir>/sleqp_b200 after `make -C sleqp_b200/csrc`).

find_path(B200_INCLUDE_DIRS
  NAMES sleqp_b200.h
  PATHS ${B200_DIR} $ENV{B200_DIR} ${INCLUDE_INSTALL_DIR}
  PATH_SUFFIXES include)

find_library(B200_LIBRARY sleqp_b200
  PATHS ${B200_DIR} $ENV{B200_DIR} ${LIB_INSTALL_DIR}
  PATH_SUFFIXES sleqp_b200 lib)

set(B200_LIBRARIES "${B200_LIBRARY}")

if(B200_INCLUDE_DIRS AND EXISTS "${B200_INCLUDE_DIRS}/sleqp_b200.h")
  file(STRINGS "${B200_INCLUDE_DIRS}/sleqp_b200.h" _B200_VERSION_LINE
    REGEX "^#define B200_VERSION_STRING")
  string(REGEX REPLACE ".*\"(.*)\".*" "\\1" B200_VERSION "${_B200_VERSION_LINE}")
endif()

include(FindPackageHandleStandardArgs)

find_package_handle_standard_args(B200
  REQUIRED_VARS B200_INCLUDE_DIRS B200_LIBRARIES
  VERSION_VAR B200_VERSION)

mark_as_advanced(B200_INCLUDE_DIRS
  B200_LIBRARIES)
