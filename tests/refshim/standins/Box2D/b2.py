"""Box2D.b2 names imported by the reference (mrp00:6, mrp02:6)."""
from . import (b2Vec2 as vec2, polygonShape, circleShape, fixtureDef, contactListener, staticBody, dynamicBody,  # noqa: F401
               dot)
