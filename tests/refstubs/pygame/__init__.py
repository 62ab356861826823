"""TEST-ONLY stand-in for pygame 2.6.1 (requirements.txt:1 of the reference).

pygame is not installable in this sandbox (no network).  The reference's
environment step only uses a sliver of it for *arithmetic*:

  * ``Rect``      - int truncation of float args + strict-overlap ``colliderect``
                    (call sites: env/sprite.py:62,68-73,90-91,137-148,614-630,
                    env/bots/simple_bots.py:219-237,317-327)
  * ``Vector2``   - double precision 2-vector: + - * dot normalize rotate
                    (env/sprite.py:317-324, env/util.py:17,29-42)
  * ``time.get_ticks`` - wall clock, replaced by a virtual clock in the harness
                    (env/sprite.py:585)

The semantics below restate pygame 2.6.1's C sources (src_c/rect.c,
src_c/math.c) as published; they are what the oracle and the CUDA kernels are
pinned against.  Everything else (display, font, event, image, key constants)
is a no-op so that the reference modules import and run headless.

This file is test infrastructure: nothing under alphatank_b200/ imports it.
"""
import math as _math

QUIT = 256
KEYDOWN = 768
SRCALPHA = 0x00010000

# key constants referenced at import time by configs/config_basic.py and
# configs/config_teams.py (values are irrelevant: no keyboard in the harness)
for _i, _n in enumerate(
        "a b c d e f g h i j k l m n o p q r s t u v w x y z".split()):
    globals()["K_" + _n] = 97 + _i
K_SPACE = 32
K_UP, K_DOWN, K_RIGHT, K_LEFT = 1073741906, 1073741905, 1073741903, 1073741904
K_RETURN, K_ESCAPE, K_RSHIFT, K_LSHIFT, K_RCTRL, K_LCTRL = 13, 27, 1, 2, 3, 4
for _i in range(10):
    globals()["K_%d" % _i] = 48 + _i
    globals()["K_KP%d" % _i] = 1073741912 + _i
K_KP_ENTER, K_KP_PLUS, K_KP_MINUS, K_KP_PERIOD = 1073741912 + 20, 1073741912 + 21, 1073741912 + 22, 1073741912 + 23
K_COMMA, K_PERIOD, K_SLASH, K_SEMICOLON, K_QUOTE = 44, 46, 47, 59, 39
K_LEFTBRACKET, K_RIGHTBRACKET, K_BACKSLASH, K_MINUS, K_EQUALS = 91, 93, 92, 45, 61
K_TAB, K_BACKSPACE, K_RALT, K_LALT = 9, 8, 5, 6


def _trunc(v):
    """C ``(int)double`` as done by pg_IntFromObj for float arguments."""
    return int(v)  # Python int() truncates toward zero, like the C cast


class Rect:
    """pygame.Rect: four C ints; float arguments are truncated toward zero."""
    __slots__ = ("x", "y", "w", "h")

    def __init__(self, *args):
        if len(args) == 1:
            args = tuple(args[0])
        if len(args) == 2:
            (x, y), (w, h) = args
        else:
            x, y, w, h = args
        self.x, self.y, self.w, self.h = _trunc(x), _trunc(y), _trunc(w), _trunc(h)

    # src_c/rect.c: _pg_do_rects_intersect (non-normalised fast path; all
    # rects built by the reference have w,h > 0)
    def colliderect(self, o):
        if self.w == 0 or self.h == 0 or o.w == 0 or o.h == 0:
            return False
        return (self.x < o.x + o.w and self.y < o.y + o.h and
                self.x + self.w > o.x and self.y + self.h > o.y)

    width = property(lambda s: s.w)
    height = property(lambda s: s.h)
    left = property(lambda s: s.x)
    top = property(lambda s: s.y)
    right = property(lambda s: s.x + s.w)
    bottom = property(lambda s: s.y + s.h)
    topleft = property(lambda s: (s.x, s.y))
    topright = property(lambda s: (s.x + s.w, s.y))
    bottomright = property(lambda s: (s.x + s.w, s.y + s.h))
    bottomleft = property(lambda s: (s.x, s.y + s.h))
    center = property(lambda s: (s.x + s.w // 2, s.y + s.h // 2))

    def __repr__(self):
        return "<rect(%d, %d, %d, %d)>" % (self.x, self.y, self.w, self.h)


class Vector2:
    """pygame.math.Vector2 (src_c/math.c): two C doubles."""
    __slots__ = ("x", "y")

    def __init__(self, x=0.0, y=None):
        if y is None:
            if isinstance(x, Vector2):
                x, y = x.x, x.y
            elif isinstance(x, (tuple, list)):
                x, y = x
            else:
                y = x
        self.x, self.y = float(x), float(y)

    def __add__(self, o):
        o = o if isinstance(o, Vector2) else Vector2(o)
        return Vector2(self.x + o.x, self.y + o.y)
    __radd__ = __add__

    def __sub__(self, o):
        o = o if isinstance(o, Vector2) else Vector2(o)
        return Vector2(self.x - o.x, self.y - o.y)

    def __rsub__(self, o):
        return Vector2(o).__sub__(self)

    def __mul__(self, k):
        if isinstance(k, Vector2):
            return self.dot(k)
        return Vector2(self.x * k, self.y * k)
    __rmul__ = __mul__

    def __neg__(self):
        return Vector2(-self.x, -self.y)

    def __iter__(self):
        yield self.x
        yield self.y

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __len__(self):
        return 2

    # _scalar_product: ret = 0; ret += a[i]*b[i]   (plain C, no contraction in
    # the manylinux x86-64 wheels, which target baseline SSE2)
    def dot(self, o):
        o = o if isinstance(o, Vector2) else Vector2(o)
        return (0.0 + self.x * o.x) + self.y * o.y

    def length(self):
        return _math.sqrt(self.dot(self))

    def normalize(self):
        ln = _math.sqrt(self.dot(self))
        if ln == 0:
            raise ValueError("Can't normalize Vector of length Zero")
        return Vector2(self.x / ln, self.y / ln)

    # _vector2_rotate_helper(dst, src, angle_deg, epsilon=1e-6)
    def rotate(self, angle):
        eps = 1e-6
        a = _math.fmod(float(angle), 360.0)
        if a < 0:
            a += 360.0
        if _math.fmod(a + eps, 90.0) < 2 * eps:
            q = int((a + eps) / 90.0)
            if q in (0, 4):
                return Vector2(self.x, self.y)
            if q == 1:
                return Vector2(-self.y, self.x)
            if q == 2:
                return Vector2(-self.x, -self.y)
            return Vector2(self.y, -self.x)
        r = a * _math.pi / 180.0          # DEG2RAD(angle) = angle * M_PI / 180.0
        s, c = _math.sin(r), _math.cos(r)
        return Vector2(c * self.x - s * self.y, s * self.x + c * self.y)

    def __repr__(self):
        return "<Vector2(%r, %r)>" % (self.x, self.y)


class _math_ns:
    Vector2 = Vector2


math = _math_ns


class _Time:
    """Virtual clock. The harness sets ``now_ms`` before every env.step()."""
    now_ms = 0

    def get_ticks(self):
        return self.now_ms

    def wait(self, ms):
        return ms

    class Clock:
        def tick(self, fps=0):
            return 0

        def get_fps(self):
            return 0.0


time = _Time()


class _Keys:
    def __getitem__(self, k):
        return False


class _Key:
    def get_pressed(self):
        return _Keys()


key = _Key()


class _Event:
    def get(self):
        return []


event = _Event()


class _Surface:
    def __init__(self, *a, **k):
        pass

    def fill(self, *a, **k):
        pass

    def blit(self, *a, **k):
        pass

    def get_rect(self, **k):
        return Rect(0, 0, 1, 1)


Surface = _Surface


class _Font:
    def init(self):
        pass

    def Font(self, *a, **k):
        return self

    SysFont = Font

    def render(self, *a, **k):
        return _Surface()


font = _Font()


class _Image:
    def fromstring(self, data, size, mode):
        return _Surface()

    frombuffer = fromstring


image = _Image()


class _Display:
    def set_mode(self, *a, **k):
        time_started[0] = True
        return _Surface()

    def init(self):
        pass

    def quit(self):
        pass

    def flip(self):
        pass

    def get_surface(self):
        return _Surface()

    def set_caption(self, *a):
        pass


time_started = [False]
display = _Display()


class _Draw:
    def __getattr__(self, name):
        return lambda *a, **k: None


draw = _Draw()
transform = _Draw()


def init():
    return (0, 0)


def quit():
    pass
