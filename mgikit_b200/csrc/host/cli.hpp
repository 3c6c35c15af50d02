// `mgikit demultiplex` command line — same flags, short forms and defaults as
// the reference (/root/reference/src/main.rs:93-399, listed in SURVEY.md A.1).
#pragma once

#include "pipeline.hpp"

namespace mgikit {

// Parses argv, runs the command on the given hot-path backend, returns the
// process exit status (0 ok, 2 usage error, 101 fatal error like a Rust panic).
int cli_main(int argc, char** argv, const Backend& be);

}  // namespace mgikit
