"""graphlearn_b200.sampler - see the package README / DESIGN.md for the layer map."""
