// Only needed when the wrappers are built against the GNU Radio stand-in headers (oracle/shim): the stand-in
// logger macros reference this flag.  Not part of a real GNU Radio build.
extern "C" { int oracle_shim_quiet = 0; }
