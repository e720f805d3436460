"""Hardware harness scripts that compare the engine with the CPU oracle (run by tools/*.sh through gpurun)."""
