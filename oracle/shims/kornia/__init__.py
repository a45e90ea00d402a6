"""Stand-in for kornia 0.7.2 (not installed): only the four functions the reference's hot path calls.
Used ONLY by oracle/ref_loader.py to execute the reference's own files in the build container."""
