"""Drop-in command-line tools (same names, flags, schemas and exit codes as the reference's scripts/)."""
