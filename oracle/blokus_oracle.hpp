// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// CPU restatement (C++17) of the reference's `blokus` crate, keeping the reference's data
// structures and control flow (nested bool vectors, per-player piece lists that are deep-cloned,
// a tile -> set-of-placements map narrowed by set intersection).  Only tests/, bench.py's
// cpu_baseline / `--impl reference` leg and __graft_entry__.smoke() may use it.
//
// Parity status: the reference's own unit tests pin ONLY the piece tables
// (blokus/src/pieces.rs:225-301) and one is_valid_move case (blokus/src/board.rs:220-225); this
// file passes all of them (tests/test_oracle_reference_kats.py).  Nothing in the reference pins
// game.rs (apply / advance_player / scoring / planes): for those, "parity unpinned" — the
// restatement is checked against the derived known answers of SURVEY.md Appendix C and against
// an independently written second restatement (oracle/py_restatement.py, tests/test_py_restatement.py).  No Rust toolchain exists in this image,
// so the reference itself cannot be run here.
//
// Canonicalisation: where the reference iterates a HashMap/HashSet the
// oracle iterates in ascending key order (std::map / std::set == BTreeMap / BTreeSet).
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <unordered_set>
#include <utility>
#include <vector>

namespace orc {

constexpr size_t BOARD_SIZE = 20;            // blokus/src/board.rs:9
constexpr int TOTAL_TILES = 89;              // blokus/src/board.rs:10
constexpr size_t BOARD_SPACES = 400;         // blokus/src/game.rs:8
constexpr size_t NUM_PLAYERS = 4;            // blokus/src/game.rs:9
constexpr size_t NUM_PIECE_TYPES = 21;       // blokus/src/pieces.rs:30

using Shape = std::vector<std::vector<bool>>;

// The 21 shapes of blokus/src/pieces.rs:124-147, in PIECE_TYPES order (pieces.rs:30-52),
// written as rows of 'X' (filled) / '.' (blank).
inline const std::vector<std::vector<std::string>>& piece_art() {
    static const std::vector<std::vector<std::string>> art = {
        {"X"},                      // One
        {"XX"},                     // Two
        {"XX", ".X"},               // Right
        {"XXX"},                    // Three
        {"XXXX"},                   // Four
        {"XX", "X.", "X."},         // ShortL
        {"XXX", ".X."},             // Triangle
        {"XX", "XX"},               // Square
        {"XX.", ".XX"},             // ShortStep
        {"XXXXX"},                  // Five
        {"XXXX", "X..."},           // LongL
        {"XXX.", "..XX"},           // LongStep
        {"XXX", "XX."},             // SquarePlus
        {"XXX", "X..", "X.."},      // LongRight
        {"XX.", ".XX", "..X"},      // Steps
        {"XX.", ".X.", ".XX"},      // Z
        {"XXX", "X.X"},             // Hump
        {"XXXX", ".X.."},           // LongWithSide
        {".X.", "XXX", ".X."},      // Plus
        {".X.", "XXX", "X.."},      // Crazy
        {"XXX", ".X.", ".X."},      // T
    };
    return art;
}

inline Shape shape_from_art(const std::vector<std::string>& rows) {
    Shape s;
    for (const auto& r : rows) {
        std::vector<bool> row;
        for (char c : r) row.push_back(c == 'X');
        s.push_back(row);
    }
    return s;
}

// blokus/src/pieces.rs:57-103
struct PieceVariant {
    std::vector<size_t> offsets;
    std::vector<bool> variant;
    size_t width;
    Shape shape;

    explicit PieceVariant(const Shape& shp) : width(shp[0].size()), shape(shp) {
        // pieces.rs:70-84 — every row but the last is padded out to the board stride
        for (size_t i = 0; i < shp.size(); ++i) {
            for (bool sq : shp[i]) variant.push_back(sq);
            if (i == shp.size() - 1) continue;
            for (size_t k = 0; k < BOARD_SIZE - shp[i].size(); ++k) variant.push_back(false);
        }
        // pieces.rs:87-91
        for (size_t i = 0; i < variant.size(); ++i)
            if (variant[i]) offsets.push_back(i);
    }
    Shape get_shape() const { return shape; }                         // pieces.rs:100-102
    bool operator==(const PieceVariant& o) const { return variant == o.variant; }  // pieces.rs:105-109
};

// blokus/src/pieces.rs:112-216
struct Piece {
    size_t id;
    Shape shape;
    uint32_t points;
    std::vector<PieceVariant> variants;

    // pieces.rs:159-170
    static Shape rotate(const Shape& shape) {
        Shape out;
        for (size_t i = 0; i < shape[0].size(); ++i) {
            std::vector<bool> row;
            for (size_t j = shape.size(); j-- > 0;) row.push_back(shape[j][i]);
            out.push_back(row);
        }
        return out;
    }
    // pieces.rs:173-183
    static Shape flip(const Shape& shape) {
        Shape out;
        for (const auto& row : shape) out.emplace_back(row.rbegin(), row.rend());
        return out;
    }
    // pieces.rs:185-209
    static std::vector<PieceVariant> gen_variants(const Shape& shape) {
        std::vector<PieceVariant> variants;
        auto push_unique = [&variants](const Shape& s) {
            PieceVariant v(s);
            for (const auto& have : variants)
                if (have == v) return;
            variants.push_back(v);
        };
        Shape cur = shape;
        for (int k = 0; k < 4; ++k) { push_unique(cur); cur = rotate(cur); }
        cur = flip(shape);
        for (int k = 0; k < 4; ++k) { push_unique(cur); cur = rotate(cur); }
        return variants;
    }
    // pieces.rs:124-156
    explicit Piece(size_t piece_type)
        : id(piece_type), shape(shape_from_art(piece_art().at(piece_type))), points(0),
          variants(gen_variants(shape)) {
        for (const auto& row : shape)
            for (bool b : row) points += b ? 1u : 0u;
    }
    bool operator==(const Piece& o) const { return shape == o.shape; }  // pieces.rs:212-216
};

// blokus/src/board.rs:18-206
struct Board {
    std::array<uint8_t, BOARD_SPACES> board{};
    std::array<std::vector<Piece>, 4> pieces;
    std::array<std::unordered_set<size_t>, 4> anchors;

    // board.rs:26-60
    Board() {
        std::vector<Piece> all;
        for (size_t t = 0; t < NUM_PIECE_TYPES; ++t) all.emplace_back(t);
        for (auto& p : pieces) p = all;
        const size_t starts[4] = {0, BOARD_SIZE - 1, BOARD_SIZE * BOARD_SIZE - 1,
                                  BOARD_SIZE * (BOARD_SIZE - 1)};
        for (size_t i = 0; i < 4; ++i) anchors[i].insert(starts[i]);
        board.fill(0);
    }

    // board.rs:62-92
    bool is_valid_move(size_t player, const PieceVariant& pv, size_t offset) const {
        const auto& variant = pv.variant;
        if (offset + variant.size() > board.size()) return false;
        if (offset % BOARD_SIZE + pv.width > BOARD_SIZE) return false;
        const uint8_t player_restricted = uint8_t(1u << (player + 4));
        bool on_blanks = true;
        for (size_t i = 0; i < variant.size(); ++i)
            if (variant[i] && (board[offset + i] & player_restricted)) { on_blanks = false; break; }
        bool on_anchor = false;
        for (size_t i : pv.offsets)
            if (anchors[player].count(offset + i)) { on_anchor = true; break; }
        return on_blanks && on_anchor;
    }

    // board.rs:95-141
    void place_tile(size_t tile, size_t player) {
        board[tile] = uint8_t(0xF0u | (player + 1));
        const uint8_t player_restricted = uint8_t(1u << (player + 4));
        const std::pair<bool, long> neighbors[4] = {
            {tile % BOARD_SIZE > 0, -1},
            {tile % BOARD_SIZE < BOARD_SIZE - 1, 1},
            {tile >= BOARD_SIZE, -long(BOARD_SIZE)},
            {tile < BOARD_SIZE * (BOARD_SIZE - 1), long(BOARD_SIZE)},
        };
        for (auto& a : anchors) a.erase(tile);
        for (const auto& nb : neighbors) {
            if (!nb.first) continue;
            size_t n = size_t(long(tile) + nb.second);
            board[n] |= player_restricted;
            anchors[player].erase(n);
        }
        // board.rs:11-16 CORNERS_OFFSETS
        const int corner_offsets[4] = {1 + int(BOARD_SIZE), -1 - int(BOARD_SIZE),
                                       1 - int(BOARD_SIZE), -1 + int(BOARD_SIZE)};
        for (int co : corner_offsets) {
            int corner = int(tile) + co;
            if (corner < 0 || corner >= int(BOARD_SPACES) ||
                (board[size_t(corner)] & player_restricted))
                continue;
            if (tile % BOARD_SIZE == 0 && size_t(corner) % BOARD_SIZE == BOARD_SIZE - 1) continue;
            if (tile % BOARD_SIZE == BOARD_SIZE - 1 && size_t(corner) % BOARD_SIZE == 0) continue;
            anchors[player].insert(size_t(corner));
        }
    }

    std::unordered_set<size_t> get_anchors(size_t player) const { return anchors[player]; }  // :143
    std::vector<Piece> get_pieces(size_t player) const { return pieces[player]; }            // :147
    void use_piece(size_t player, size_t piece) {                                            // :151
        pieces[player].erase(pieces[player].begin() + long(piece));
    }

    // board.rs:155-181
    std::vector<int> get_scores(const std::array<uint32_t, 4>& last_piece_lens) const {
        std::vector<int> scores(4, 0);
        for (uint8_t cell : board) {
            uint8_t player = cell & 0x0F;
            if (player != 0) scores[player - 1] += 1;
        }
        for (size_t i = 0; i < 4; ++i) {
            scores[i] -= TOTAL_TILES;
            if (pieces[i].empty()) {
                scores[i] += 15;
                if (last_piece_lens[i] == 1) scores[i] += 5;
            }
        }
        return scores;
    }
};

using Move = std::tuple<size_t, size_t, size_t>;  // (piece index in REMAINING list, variant, offset)
using TileMoves = std::map<size_t, std::set<Move>>;  // canonical stand-in for HashMap<usize,HashSet<..>>

// blokus/src/game.rs:12-44
inline void get_piece_moves(size_t piece_i, const Board& board, size_t player,
                            std::vector<Move>& moves, std::vector<std::vector<size_t>>& groups) {
    const std::vector<Piece> pieces = board.get_pieces(player);  // deep clone, as game.rs:19
    const Piece& piece = pieces[piece_i];
    const auto anchors = board.get_anchors(player);              // clone, as game.rs:20
    for (size_t anchor : anchors) {
        for (size_t var_i = 0; var_i < piece.variants.size(); ++var_i) {
            const PieceVariant& variant = piece.variants[var_i];
            for (size_t offset : variant.offsets) {
                if (offset > anchor) continue;
                size_t total_offset = anchor - offset;
                if (board.is_valid_move(player, variant, total_offset)) {
                    std::vector<size_t> tiles;
                    for (size_t j = 0; j < variant.variant.size(); ++j)
                        if (variant.variant[j]) tiles.push_back(total_offset + j);
                    groups.push_back(tiles);
                    moves.emplace_back(piece_i, var_i, total_offset);
                }
            }
        }
    }
}

// blokus/src/game.rs:47-57
inline void get_moves(const Board& board, size_t player, std::vector<Move>& moves,
                      std::vector<std::vector<size_t>>& groups) {
    const size_t n = board.get_pieces(player).size();
    for (size_t piece = 0; piece < n; ++piece) get_piece_moves(piece, board, player, moves, groups);
}

// blokus/src/game.rs:60-74
inline TileMoves get_tile_moves(const Board& board, size_t player) {
    TileMoves rep;
    std::vector<Move> moves;
    std::vector<std::vector<size_t>> groups;
    get_moves(board, player, moves, groups);
    for (size_t k = 0; k < moves.size(); ++k)
        for (size_t tile : groups[k]) rep[tile].insert(moves[k]);
    return rep;
}

using Planes = std::array<std::array<std::array<bool, 20>, 20>, 5>;

// blokus/src/game.rs:77-89
inline Planes rotate_state(const Planes& state) {
    Planes out = state;
    for (size_t i = 0; i < NUM_PLAYERS + 1; ++i)
        for (size_t j = 0; j < 20; ++j)
            for (size_t k = 0; k < 20; ++k) out[i][j][k] = state[i][k][20 - j - 1];
    return out;
}

// blokus/src/game.rs:91-312
struct Game {
    Board board;
    std::vector<std::pair<int, int>> history;  // (player, tile)
    std::array<bool, 4> eliminated{};
    size_t current_player_ = 0;
    TileMoves legal_tiles;
    std::array<uint32_t, 4> last_piece_lens{};

    // game.rs:102-114
    static Game reset() {
        Game g;
        g.legal_tiles = get_tile_moves(g.board, 0);
        return g;
    }

    // game.rs:150-194.  Returns "" on success, the reference's Err text otherwise (the state is
    // then already mutated, as in the reference).  piece_to_finish < 0 means None.
    std::string apply(size_t tile, long piece_to_finish = -1) {
        board.place_tile(tile, current_player_);
        history.emplace_back(int(current_player_), int(tile));
        auto it = legal_tiles.find(tile);
        if (it == legal_tiles.end())
            return "Invalid move - Player " + std::to_string(current_player_) + ", Tile " +
                   std::to_string(tile);
        const std::set<Move> valid_moves = it->second;
        legal_tiles.erase(it);
        const TileMoves snapshot = legal_tiles;  // game.rs:165 clones the map
        for (const auto& kv : snapshot) {
            std::set<Move> inter;
            for (const Move& m : kv.second)
                if (valid_moves.count(m)) inter.insert(m);
            if (inter.empty()) legal_tiles.erase(kv.first);
            else legal_tiles[kv.first] = inter;
        }
        if (legal_tiles.empty() || piece_to_finish >= 0) {
            size_t piece = piece_to_finish >= 0 ? size_t(piece_to_finish)
                                                : std::get<0>(*valid_moves.begin());
            std::vector<Piece> remaining = board.get_pieces(current_player_);
            last_piece_lens[current_player_] = remaining.at(piece).points;
            board.use_piece(current_player_, piece);
            advance_player();
        }
        return "";
    }

    // game.rs:116-144
    std::pair<Game, std::string> place_piece(size_t p, size_t v, size_t o) const {
        Game ns = *this;
        const size_t player = current_player_;
        const PieceVariant piece = get_piece(player, p, v);
        if (!ns.board.is_valid_move(player, piece, o)) return {ns, "Invalid move"};
        const size_t last = piece.offsets.empty() ? 0 : piece.offsets.size() - 1;
        for (size_t i = 0; i < piece.offsets.size(); ++i) {
            std::string e = ns.apply(o + piece.offsets[i], i == last ? long(p) : -1);
            if (!e.empty()) return {ns, e};
        }
        return {ns, ""};
    }

    const std::array<uint8_t, 400>& get_board() const { return board.board; }  // game.rs:196

    // game.rs:203-223
    size_t advance_player() {
        if (is_terminal()) return current_player_;
        current_player_ = (current_player_ + 1) % NUM_PLAYERS;
        legal_tiles = get_tile_moves(board, current_player_);
        if (eliminated[current_player_]) {
            advance_player();
        } else if (legal_tiles.empty()) {
            eliminated[current_player_] = true;
            advance_player();
        }
        return current_player_;
    }

    size_t current_player() const { return current_player_; }                       // :225
    std::vector<Piece> get_current_player_pieces() const { return board.get_pieces(current_player_); }
    PieceVariant get_piece(size_t player, size_t piece, size_t variant) const {     // :234
        return board.get_pieces(player).at(piece).variants.at(variant);
    }
    std::set<size_t> get_current_anchors() const {                                  // :238
        auto a = board.get_anchors(current_player_);
        return std::set<size_t>(a.begin(), a.end());
    }
    std::vector<size_t> get_legal_tiles() const {                                   // :242 (ascending)
        std::vector<size_t> out;
        for (const auto& kv : legal_tiles) out.push_back(kv.first);
        return out;
    }
    std::vector<int> get_score() const { return board.get_scores(last_piece_lens); }  // :247

    // game.rs:252-272
    std::vector<float> get_payoff() const {
        std::vector<int> scores = board.get_scores(last_piece_lens);
        std::vector<float> payoff(4, 0.0f);
        std::vector<size_t> indices;
        int highest = scores[0];
        for (size_t i = 0; i < scores.size(); ++i) {
            if (scores[i] == highest) indices.push_back(i);
            else if (scores[i] > highest) { indices.clear(); indices.push_back(i); highest = scores[i]; }
        }
        for (size_t i : indices) payoff[i] = 1.0f / float(indices.size());
        return payoff;
    }

    bool is_terminal() const {                                                      // :275
        for (bool e : eliminated) if (!e) return false;
        return true;
    }
    bool is_player_active(size_t player) const { return !eliminated[player]; }     // :279

    // game.rs:283-311
    Planes get_board_state() const {
        Planes st{};
        for (size_t i = 0; i < BOARD_SPACES; ++i) {
            size_t player = board.board[i] & 0x0F;
            if (player != 0) {
                size_t plane = (4 + (player - 1) - current_player_) % 4;
                st[plane][i / 20][i % 20] = true;
            }
        }
        for (size_t tile : get_legal_tiles()) st[4][tile / 20][tile % 20] = true;
        for (size_t k = 0; k < current_player_; ++k) st = rotate_state(st);
        return st;
    }
};

}  // namespace orc
