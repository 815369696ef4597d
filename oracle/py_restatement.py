"""SECOND, INDEPENDENT CPU RESTATEMENT — TEST INFRASTRUCTURE ONLY (small cases; pure-Python loops).

The reference's rules engine is Rust and cannot be compiled in this image (no rustc/cargo), so the C++ oracle
(oracle/blokus_oracle.hpp) is pinned by the reference's own tests only for the piece tables.  This module is a
second restatement of the same files, written separately from the C++ one and in a different style — Python dicts
and sets standing in for the reference's HashMap/HashSet, statement for statement — so that tests/ can triangulate:
two independent readings of blokus/src/{pieces,board,game}.rs and of self_play/src/{node,simulation}.rs must agree
on every ply of seeded games (tests/test_py_restatement.py).  It shares nothing with the oracle except the published
RNG-free policies used to drive the games (always the smallest / largest legal tile, or a caller-supplied index).

Only tests/ may import this.  Citations are into /root/reference.
"""
from __future__ import annotations

import ctypes
import math
import struct

BOARD_SIZE = 20                      # board.rs:9
TOTAL_TILES = 89                     # board.rs:10
CORNERS_OFFSETS = (1 + BOARD_SIZE, -1 - BOARD_SIZE, 1 - BOARD_SIZE, -1 + BOARD_SIZE)   # board.rs:11-16

# pieces.rs:124-147, in PIECE_TYPES order (pieces.rs:30-52); 'X' = true
_SHAPES = [
    ["X"],
    ["XX"],
    ["XX", ".X"],
    ["XXX"],
    ["XXXX"],
    ["XX", "X.", "X."],
    ["XXX", ".X."],
    ["XX", "XX"],
    ["XX.", ".XX"],
    ["XXXXX"],
    ["XXXX", "X..."],
    ["XXX.", "..XX"],
    ["XXX", "XX."],
    ["XXX", "X..", "X.."],
    ["XX.", ".XX", "..X"],
    ["XX.", ".X.", ".XX"],
    ["XXX", "X.X"],
    ["XXXX", ".X.."],
    [".X.", "XXX", ".X."],
    [".X.", "XXX", "X.."],
    ["XXX", ".X.", ".X."],
]


def _shape(rows):
    return [[c == "X" for c in r] for r in rows]


class PieceVariant:                  # pieces.rs:58-98
    def __init__(self, shape):
        variant = []
        for i, row in enumerate(shape):
            variant.extend(row)
            if i == len(shape) - 1:
                continue
            variant.extend([False] * (BOARD_SIZE - len(row)))
        self.variant = variant
        self.offsets = [i for i, sq in enumerate(variant) if sq]
        self.width = len(shape[0])
        self.shape = shape

    def __eq__(self, other):         # pieces.rs:105-109
        return self.variant == other.variant


def rotate(shape):                   # pieces.rs:159-170
    return [[shape[j][i] for j in reversed(range(len(shape)))] for i in range(len(shape[0]))]


def flip(shape):                     # pieces.rs:173-183
    return [list(reversed(row)) for row in shape]


def gen_variants(shape):             # pieces.rs:185-209
    variants = []
    vs = [list(r) for r in shape]
    for _ in range(4):
        nv = PieceVariant([list(r) for r in vs])
        if nv not in variants:
            variants.append(nv)
        vs = rotate(vs)
    vs = flip(shape)
    for _ in range(4):
        nv = PieceVariant([list(r) for r in vs])
        if nv not in variants:
            variants.append(nv)
        vs = rotate(vs)
    return variants


class Piece:                         # pieces.rs:112-156
    def __init__(self, pid):
        self.id = pid
        self.shape = _shape(_SHAPES[pid])
        self.points = sum(sum(1 for x in r if x) for r in self.shape)
        self.variants = gen_variants(self.shape)


_PIECES = [Piece(i) for i in range(21)]      # immutable here, so shared instead of cloned


class Board:                         # board.rs:18-181
    def __init__(self):
        self.board = [0] * (BOARD_SIZE * BOARD_SIZE)
        self.pieces = [list(_PIECES) for _ in range(4)]
        self.anchors = [{0}, {BOARD_SIZE - 1}, {BOARD_SIZE * BOARD_SIZE - 1}, {BOARD_SIZE * (BOARD_SIZE - 1)}]

    def clone(self):
        b = Board.__new__(Board)
        b.board = list(self.board)
        b.pieces = [list(p) for p in self.pieces]
        b.anchors = [set(a) for a in self.anchors]
        return b

    def is_valid_move(self, player, pv, offset):         # board.rs:62-92
        variant = pv.variant
        if offset + len(variant) > len(self.board):
            return False
        elif offset % BOARD_SIZE + pv.width > BOARD_SIZE:
            return False
        restricted = 1 << (player + 4)
        on_blanks = all(not (b and (a & restricted)) for a, b in zip(self.board[offset:offset + len(variant)], variant))
        on_anchor = any((offset + i) in self.anchors[player] for i in pv.offsets)
        return on_blanks and on_anchor

    def place_tile(self, tile, player):                  # board.rs:95-141
        self.board[tile] = 0b1111_0000 | (player + 1)
        restricted = 1 << (player + 4)
        neighbors = [
            (tile % BOARD_SIZE > 0, -1),
            (tile % BOARD_SIZE < BOARD_SIZE - 1, 1),
            (tile >= BOARD_SIZE, -BOARD_SIZE),
            (tile < BOARD_SIZE * (BOARD_SIZE - 1), BOARD_SIZE),
        ]
        for i in range(4):
            self.anchors[i].discard(tile)
        for in_bounds, off in neighbors:
            if in_bounds:
                nb = tile + off
                self.board[nb] |= restricted
                self.anchors[player].discard(nb)
        for co in CORNERS_OFFSETS:
            corner = tile + co
            if corner < 0 or corner >= BOARD_SIZE * BOARD_SIZE or (self.board[corner] & restricted):
                continue
            if tile % BOARD_SIZE == 0 and corner % BOARD_SIZE == BOARD_SIZE - 1:
                continue
            if tile % BOARD_SIZE == BOARD_SIZE - 1 and corner % BOARD_SIZE == 0:
                continue
            self.anchors[player].add(corner)

    def get_scores(self, last_piece_lens):               # board.rs:155-181
        scores = [0, 0, 0, 0]
        for cell in self.board:
            p = cell & 0b1111
            if p != 0:
                scores[p - 1] += 1
        for i in range(4):
            scores[i] -= TOTAL_TILES
            if len(self.pieces[i]) == 0:
                scores[i] += 15
                if last_piece_lens[i] == 1:
                    scores[i] += 5
        return scores


def get_tile_moves(board, player):                       # game.rs:12-74 (get_piece_moves, get_moves folded in)
    tile_rep = {}
    for piece_i, piece in enumerate(board.pieces[player]):
        for anchor in board.anchors[player]:
            for var_i, pv in enumerate(piece.variants):
                for off in pv.offsets:
                    if off > anchor:
                        continue
                    total = anchor - off
                    if board.is_valid_move(player, pv, total):
                        for j, sq in enumerate(pv.variant):
                            if sq:
                                tile_rep.setdefault(total + j, set()).add((piece_i, var_i, total))
    return tile_rep


def rotate_state(state):                                 # game.rs:77-89
    D = BOARD_SIZE
    return [[[state[i][k][D - j - 1] for k in range(D)] for j in range(D)] for i in range(5)]


class Game:                                              # game.rs:91-311
    def __init__(self):
        self.board = Board()
        self.history = []
        self.eliminated = [False] * 4
        self.current_player = 0
        self.legal_tiles = get_tile_moves(self.board, 0)
        self.last_piece_lens = [0] * 4

    def clone(self):
        g = Game.__new__(Game)
        g.board = self.board.clone()
        g.history = list(self.history)
        g.eliminated = list(self.eliminated)
        g.current_player = self.current_player
        g.legal_tiles = {t: set(m) for t, m in self.legal_tiles.items()}
        g.last_piece_lens = list(self.last_piece_lens)
        return g

    def apply(self, tile, piece_to_finish=None):         # game.rs:150-194
        self.board.place_tile(tile, self.current_player)
        self.history.append((self.current_player, tile))
        valid_moves = self.legal_tiles.pop(tile, None)
        if valid_moves is None:
            raise ValueError(f"Invalid move - Player {self.current_player}, Tile {tile}")
        for t, move_set in list(self.legal_tiles.items()):
            inter = move_set & valid_moves
            if inter:
                self.legal_tiles[t] = inter
            else:
                del self.legal_tiles[t]
        if len(self.legal_tiles) == 0 or piece_to_finish is not None:
            piece = piece_to_finish if piece_to_finish is not None else next(iter(valid_moves))[0]
            self.last_piece_lens[self.current_player] = self.board.pieces[self.current_player][piece].points
            del self.board.pieces[self.current_player][piece]                                   # use_piece
            self.advance_player()

    def place_piece(self, p, v, o):                      # game.rs:116-144
        ns = self.clone()
        pv = self.board.pieces[self.current_player][p].variants[v]
        if not ns.board.is_valid_move(self.current_player, pv, o):
            raise ValueError("Invalid move")
        last = max(len(pv.offsets) - 1, 0)
        for i, off in enumerate(pv.offsets):
            ns.apply(o + off, p if i == last else None)
        return ns

    def advance_player(self):                            # game.rs:203-223
        if self.is_terminal():
            return self.current_player
        self.current_player = (self.current_player + 1) % 4
        self.legal_tiles = get_tile_moves(self.board, self.current_player)
        if self.eliminated[self.current_player]:
            self.advance_player()
        elif len(self.legal_tiles) == 0:
            self.eliminated[self.current_player] = True
            self.advance_player()
        return self.current_player

    def get_legal_tiles(self):                           # game.rs:242-244 (sorted: the order is arbitrary there)
        return sorted(self.legal_tiles)

    def get_score(self):                                 # game.rs:247-249
        return self.board.get_scores(self.last_piece_lens)

    def get_payoff(self):                                # game.rs:252-272
        scores = self.get_score()
        payoff = [0.0] * 4
        indices = []
        highest = scores[0]
        for i, sc in enumerate(scores):
            if sc == highest:
                indices.append(i)
            elif sc > highest:
                indices = [i]
                highest = sc
        for i in indices:
            payoff[i] = f32(1.0 / len(indices))
        return payoff

    def is_terminal(self):                               # game.rs:275-277
        return all(self.eliminated)

    def get_board_state(self):                           # game.rs:283-311
        D = BOARD_SIZE
        st = [[[False] * D for _ in range(D)] for _ in range(5)]
        for i, cell in enumerate(self.board.board):
            p = cell & 0b1111
            if p != 0:
                st[(4 + (p - 1) - self.current_player) % 4][i // D][i % D] = True
        for t in self.legal_tiles:
            st[4][t // D][t % D] = True
        for _ in range(self.current_player):
            st = rotate_state(st)
        return st


# ---- self_play crate: node.rs:8-41, simulation.rs:37-231 (noise-free, greedy form) ------------------------------
# f32 arithmetic is done by rounding every intermediate to binary32; ln/sqrt/exp come from the platform libm, as
# Rust's f32::ln / sqrt / exp do (LLVM intrinsics -> libm).

_libm = ctypes.CDLL("libm.so.6")
_libm.logf.restype = ctypes.c_float
_libm.logf.argtypes = [ctypes.c_float]
_libm.sqrtf.restype = ctypes.c_float
_libm.sqrtf.argtypes = [ctypes.c_float]


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


class Node:                                              # node.rs:8-41
    def __init__(self, prior):
        self.children = {}
        self.to_play = 0
        self.value_sum = 0.0
        self.visits = 0
        self.prior = prior

    def is_expanded(self):
        return len(self.children) > 0

    def value(self):
        return 0.0 if self.visits == 0 else f32(self.value_sum / self.visits)


def rotate_policy(state):                                # simulation.rs:25-34
    rot = [0.0] * 400
    for i in range(20):
        for j in range(20):
            rot[j * 20 + (19 - i)] = state[i * 20 + j]
    return rot


def evaluate(node, game, evaluator, exp_f32):            # simulation.rs:37-83
    if game.is_terminal():
        return game.get_payoff()
    policy, value = evaluator(game.get_board_state())
    cur = game.current_player
    for _ in range(cur):
        policy = rotate_policy(policy)
    value = list(value)
    for _ in range(cur):                                 # value.rotate_right(cur)
        value = [value[-1]] + value[:-1]
    node.to_play = cur
    legal = game.get_legal_tiles()                       # canonical order: ascending
    exps = []
    total = 0.0
    for t in legal:
        if policy[t] > 0.0:
            e = exp_f32(policy[t])
            exps.append((t, e))
            total = f32(total + e)
    for t, e in exps:
        node.children[t] = Node(f32(e / total))
    return value


def ucb_score(parent, child, c_base, c_init):            # simulation.rs:88-98
    pv = f32(float(parent.visits))
    a = f32(f32(pv + c_base) + 1.0)
    a = f32(a / c_base)
    c = f32(_libm.logf(a) + c_init)
    c = f32(c * _libm.sqrtf(pv))
    c = f32(c / f32(float(child.visits) + 1.0))
    prior_score = f32(c * child.prior)
    return f32(prior_score + child.value())


def select_child(node, c_base, c_init):                  # simulation.rs:135-147 (ascending order; `>=`: last maximum)
    best_score, best_action = 0.0, 0
    for action in sorted(node.children):
        sc = ucb_score(node, node.children[action], c_base, c_init)
        if sc >= best_score:
            best_score, best_action = sc, action
    return best_action


def mcts(game, sims, c_base, c_init, evaluator, exp_f32):    # simulation.rs:174-231 without noise (fraction 0)
    root = Node(0.0)
    evaluate(root, game, evaluator, exp_f32)
    for _ in range(sims):
        root.visits += 1
        node = root
        scratch = game.clone()
        path = []
        while node.is_expanded():
            action = select_child(node, c_base, c_init)
            node = node.children[action]
            scratch.apply(action, None)
            path.append(node)
        values = evaluate(node, scratch, evaluator, exp_f32)
        for n in path:                                   # backpropagate, simulation.rs:164-171
            n.visits += 1
            n.value_sum = f32(n.value_sum + values[n.to_play])
    return root


def best_action_by_visits(root):                         # select_action's greedy branch, simulation.rs:159 (last maximum)
    best, best_v = None, -1
    for a in sorted(root.children):
        if root.children[a].visits >= best_v:
            best, best_v = a, root.children[a].visits
    return best


def exp_f32_default(x):
    """f32 exp as both sides of this repo define it (DESIGN.md deviation 4): exp in f64, rounded once."""
    return f32(math.exp(x))
