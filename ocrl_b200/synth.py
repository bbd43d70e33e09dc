"""Synthetic frame pool for benchmarks and smoke tests.

Restates the *distribution* of the reference's sprite scenes -- envs/synthetic_envs/base.py:102-151
(rejection-sampled positions), randomobjs.py:15-27 (colour / shape / scale draws) with
configs/env/random-N5C4S4S2.yaml:1-11 (5 sprites, 4 colours, 4 shapes, scales 0.15 / 0.22, black
background, no agent, occlusion threshold 0.15) and the push variant configs/env/push-N3C4S1S1.yaml.
spriteworld / gym are not installable offline, so the polygons are drawn with PIL at 10x and
down-sampled (anti_aliasing=10, base.py:32-35).  Pixel parity with spriteworld is not claimed:
frames are only inputs, the oracle sees the same tensor.
"""
from __future__ import annotations

import math

import numpy as np
from PIL import Image, ImageDraw

COLORS = {"blue": (0, 0, 255), "green": (0, 128, 0), "yellow": (255, 255, 0), "red": (255, 0, 0)}
SHAPES = ("square", "triangle", "star_4", "circle")
SCALES = (0.15, 0.22)


def _polygon(shape: str) -> np.ndarray:
    """Unit-area-normalised vertex list centred at the origin."""
    if shape == "circle":
        n, star = 40, None
    elif shape == "square":
        n, star = 4, None
    elif shape == "triangle":
        n, star = 3, None
    elif shape == "star_4":
        n, star = 4, 0.4
    else:
        raise ValueError(shape)
    if star is None:
        ang = np.linspace(0, 2 * math.pi, n, endpoint=False) + (math.pi / 4 if n == 4 else math.pi / 2)
        pts = np.stack([np.cos(ang), np.sin(ang)], 1)
    else:
        ang = np.linspace(0, 2 * math.pi, 2 * n, endpoint=False) + math.pi / 2
        rad = np.where(np.arange(2 * n) % 2 == 0, 1.0, star)
        pts = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
    x, y = pts[:, 0], pts[:, 1]
    area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    return pts / math.sqrt(area)


_POLYS = {s: _polygon(s) for s in SHAPES}


def _positions(rng, scales, wall_eps=0.08, threshold=0.15, fixed=()):
    placed = [np.asarray(f, float) for f in fixed]
    out = []
    for sc in scales:
        r = sc / 2
        while True:
            xy = rng.uniform(r + wall_eps, 1 - r - wall_eps, size=2)
            if all(np.linalg.norm(xy - q) >= threshold for q in placed):
                break
        placed.append(xy)
        out.append(xy)
    return out


def render(sprites, size: int, aa: int = 10) -> np.ndarray:
    """sprites: list of (shape, rgb, scale, x, y); later sprites occlude earlier ones; y axis up."""
    big = size * aa
    img = Image.new("RGB", (big, big), (0, 0, 0))
    draw = ImageDraw.Draw(img)
    for shape, rgb, scale, x, y in sprites:
        pts = _POLYS[shape] * scale + np.array([x, y])
        draw.polygon([(float(px * big), float(py * big)) for px, py in pts], fill=tuple(rgb))
    img = img.resize((size, size), Image.LANCZOS)
    return np.flipud(np.asarray(img, dtype=np.uint8)).copy()


def random_objs_frames(count: int, size: int = 64, seed: int = 0, num_objects: int = 5) -> np.ndarray:
    """uint8 [count, size, size, 3] frames of the random-N5C4S4S2 distribution."""
    rng = np.random.RandomState(seed)
    names = list(COLORS)
    frames = np.empty((count, size, size, 3), np.uint8)
    for i in range(count):
        cols = [COLORS[names[rng.randint(len(names))]] for _ in range(num_objects)]
        shapes = [SHAPES[rng.randint(len(SHAPES))] for _ in range(num_objects)]
        scales = [SCALES[rng.randint(len(SCALES))] for _ in range(num_objects)]
        pos = _positions(rng, scales)
        frames[i] = render([(s, c, sc, p[0], p[1]) for s, c, sc, p in zip(shapes, cols, scales, pos)], size)
    return frames


def push_frames(count: int, size: int = 64, seed: int = 0) -> np.ndarray:
    """push-N3C4S1S1 (hard / sparse): 3 squares of scale 0.15 (one of them the blue target), a goal
    square in the corner (r, r) and the red circular agent at (0.5, 0.5)."""
    rng = np.random.RandomState(seed)
    names = list(COLORS)
    frames = np.empty((count, size, size, 3), np.uint8)
    sc = 0.15
    for i in range(count):
        goal = (sc / 2, sc / 2)
        agent = (0.5, 0.5)
        pos = _positions(rng, [sc] * 3, wall_eps=0.15, fixed=(goal, agent))
        cols = [COLORS["blue"]] + [COLORS[names[rng.randint(1, len(names))]] for _ in range(2)]
        sprites = [("square", c, sc, p[0], p[1]) for c, p in zip(cols, pos)]
        sprites.append(("square", COLORS["blue"], sc, goal[0], goal[1]))
        sprites.append(("circle", COLORS["red"], sc, agent[0], agent[1]))
        frames[i] = render(sprites, size)
    return frames


def to_obs(frames_u8):
    """uint8 HWC -> float32 CHW / 255 (utils/datasets.py:17), torch tensor in, torch tensor out."""
    return frames_u8.permute(0, 3, 1, 2).float() / 255.0
