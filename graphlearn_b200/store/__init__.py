"""graphlearn_b200.store - see the package README / DESIGN.md for the layer map."""
