"""Stand-in for pyglet (reference mrp00:11-12): rendering is out of scope, nothing is ever called."""
