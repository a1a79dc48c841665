"""graphlearn_b200.io - see the package README / DESIGN.md for the layer map."""
