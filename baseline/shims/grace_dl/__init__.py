"""Minimal stand-in for the un-vendored GRACE package (sands-lab/grace) so the
UNMODIFIED reference file can be imported.  Contract: SURVEY.md Appendix A."""
