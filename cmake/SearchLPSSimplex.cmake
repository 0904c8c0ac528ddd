# Goes to <sleqp>/cmake/SearchLPSSimplex.cmake: the built-in dense simplex LP backend (src/main/lp/lpi_simplex.c,
# from sleqp_b200/host/lp/) has no external dependency, so it is always "found".
# Select with -DSLEQP_LPS=Simplex after adding
#   add_lp_solver(NAME "Simplex" SOURCES lp/lpi_simplex.c)
# next to the other add_lp_solver(...) calls of cmake/SearchLPS.cmake:31-43 (done by sleqp_b200_backend.patch).
set(SIMPLEX_FOUND TRUE)
set(SIMPLEX_INCLUDE_DIRS "")
set(SIMPLEX_LIBRARIES "")
set(SIMPLEX_VERSION "0.1")
