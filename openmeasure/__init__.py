"""Import alias: `from openmeasure.sparse_sensing import ROM, SPR` -- the reference's own import line
(README.md:157 of burn-research/OpenMEASURE) -- resolves to the B200 path when this repository is on
sys.path ahead of (or instead of) the reference package.  Only the hot-path module exists here: the
reference's gpr / cokriging / utils modules are out of scope."""
