"""Problem bases and the BASELINE configurations -- the CALLERS of the hot path.

They keep the reference's behaviour (point generators, losses, CLI flags) so that
results are comparable run for run; citations in each file.  The reference's own
unchanged problem files can be used instead: these exist because
/root/reference does not travel to the GPU box."""
