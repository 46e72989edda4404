// Host-side front end of the encoder: BERT word-piece tokenisation and the sequence assembly of
// prepare_bert_sentences, natively and multi-threaded ("the step before the encoder").
//
// Replaces, per sentence, `tokenizer.tokenize(s)` + `tokenizer.convert_tokens_to_ids(...)`
// (examples/ex_aspire_consent.py:131-133 = src/learning/batchers.py:579-581) and, per document, the concatenation /
// 500-word-piece truncation / [CLS]..[SEP] wrapping / right padding / sentence-span bookkeeping of
// examples/ex_aspire_consent.py:120-173.  The encoder kernels run 15 k documents/s; the Python path manages 0.6 k and
// one batched call into the Rust tokenizer 1.4-4 k, so the tokenizer had become the pipeline's bottleneck.
//
// Scope: the BERT tokenizer the reference loads (BasicTokenizer + WordPiece: clean text, lower-case, split on
// whitespace and punctuation, greedy longest-match-first word pieces with the "##" continuation prefix, words longer
// than max_input_chars_per_word -> [UNK]; special tokens such as the literal "[SEP]" the reference appends to the title
// are matched verbatim before normalisation).  ASCII sentences take a single-pass fast path with the rules written
// out below.  Non-ASCII sentences are normalised and classified through per-code-point tables of the Basic
// Multilingual Plane that the caller reads off the tokenizer's own normaliser and pre-tokenizer
// (asp_wordpiece_set_unicode) -- accent stripping, lower-casing, CJK spacing, Unicode spaces / punctuation all come
// from there, nothing about Unicode is restated here.  What the tables cannot express (characters beyond U+FFFF,
// characters whose treatment depends on context, malformed UTF-8 -- or any non-ASCII byte when no tables were
// given) is reported back per sentence (needs_fallback) and the caller runs that sentence through the original
// tokenizer.  No CUDA in this file; it lives in the same library so the ctypes binding and the error convention are
// shared.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>
#include "common.cuh"

struct asp_wordpiece {
    std::string blob;  // owns the vocabulary text the maps' keys point into
    std::unordered_map<std::string_view, int32_t> initial, continuation;  // "abc" / "##abc" (stored without the prefix)
    std::vector<std::pair<std::string, int32_t>> specials;                // matched verbatim in the raw text
    size_t max_piece = 0;
    int32_t unk_id = 0;
    bool lower_case = true;
    // Optional per-code-point tables of the Basic Multilingual Plane (asp_wordpiece_set_unicode): what the tokenizer's
    // normaliser turns each character into (lower-casing, accent stripping, text cleaning, spaces around CJK ...), how
    // the pre-tokenizer classes each character (0 word, 1 space, 2 punctuation), and which characters must be left to
    // the original tokenizer because their treatment depends on context.
    std::atomic<bool> unicode{false};  // published with release order once the tables below are complete
    std::vector<uint32_t> norm_offsets;  // 65537
    std::string norm_blob;
    std::vector<uint8_t> out_class, fallback;  // 65536 each
};

namespace {

inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
// BasicTokenizer._clean_text drops NUL, U+FFFD and every "C*" category character except \t \n \r (treated as spaces)
inline bool is_dropped(unsigned char c) { return c == 0 || (c < 0x20 && !is_space(c)) || c == 0x7f; }
// _is_punctuation: all non-alphanumeric printable ASCII
inline bool is_punct(unsigned char c) {
    return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126);
}

// Greedy longest-match-first word pieces of one cleaned, lower-cased word (WordpieceTokenizer.tokenize).
template <typename Out>
inline void wordpiece(const asp_wordpiece& wp, const char* w, size_t n, size_t n_chars, size_t max_chars, Out& out) {
    if (n_chars > max_chars) {
        *out++ = wp.unk_id;
        return;
    }
    const auto first = out;
    size_t start = 0;
    while (start < n) {
        const auto& map = start ? wp.continuation : wp.initial;
        size_t end = std::min(n, start + wp.max_piece);
        int32_t id = -1;
        for (; end > start; --end) {
            auto it = map.find(std::string_view(w + start, end - start));
            if (it != map.end()) {
                id = it->second;
                break;
            }
        }
        if (id < 0) {  // no piece matches: the whole word is unknown
            out = first;
            *out++ = wp.unk_id;
            return;
        }
        *out++ = id;
        start = end;
    }
}

// One sentence -> ids at `out` (capacity >= its byte length: every id consumes at least one byte).  Returns the count,
// or -1 if the sentence holds a non-ASCII byte.
int64_t encode_sentence_unicode(const asp_wordpiece& wp, const char* s, size_t n, size_t max_chars, int32_t* out_base);

int64_t encode_sentence(const asp_wordpiece& wp, const char* s, size_t n, size_t max_chars, int32_t* out) {
    for (size_t i = 0; i < n; ++i)
        if ((unsigned char)s[i] >= 0x80) return wp.unicode.load(std::memory_order_acquire) ? encode_sentence_unicode(wp, s, n, max_chars, out) : -1;
    int32_t* const base = out;
    std::string word;
    word.reserve(64);
    auto flush = [&]() {
        if (!word.empty()) {
            wordpiece(wp, word.data(), word.size(), word.size(), max_chars, out);
            word.clear();
        }
    };
    size_t i = 0;
    while (i < n) {
        const unsigned char c = (unsigned char)s[i];
        if (!wp.specials.empty() && c == (unsigned char)wp.specials.front().first[0]) {
            // special tokens are cut out of the raw text before any normalisation (longest match at this position)
            const std::pair<std::string, int32_t>* hit = nullptr;
            for (const auto& sp : wp.specials)
                if (sp.first.size() <= n - i && memcmp(s + i, sp.first.data(), sp.first.size()) == 0 &&
                    (!hit || sp.first.size() > hit->first.size()))
                    hit = &sp;
            if (hit) {
                flush();
                *out++ = hit->second;
                i += hit->first.size();
                continue;
            }
        }
        ++i;
        if (is_dropped(c)) continue;
        if (is_space(c)) {
            flush();
        } else if (is_punct(c)) {
            flush();
            const char p = (char)c;
            wordpiece(wp, &p, 1, 1, max_chars, out);
        } else {
            word.push_back(wp.lower_case && c >= 'A' && c <= 'Z' ? (char)(c + 32) : (char)c);
        }
    }
    flush();
    return out - base;
}

// Decodes one UTF-8 sequence at s[i..n); returns its length (0 = malformed) and the code point.
inline int utf8_decode(const char* s, size_t i, size_t n, uint32_t& cp) {
    const unsigned char c = (unsigned char)s[i];
    if (c < 0x80) {
        cp = c;
        return 1;
    }
    const int len = (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 0;
    if (!len || i + len > n) return 0;
    cp = c & (0xFF >> (len + 1));
    for (int k = 1; k < len; ++k) {
        const unsigned char d = (unsigned char)s[i + k];
        if ((d & 0xC0) != 0x80) return 0;
        cp = (cp << 6) | (d & 0x3F);
    }
    if ((len == 2 && cp < 0x80) || (len == 3 && cp < 0x800) || (len == 4 && cp < 0x10000)) return 0;  // overlong
    return len;
}

struct IdSink {  // bounded writer: a sentence may never emit more ids than it has bytes (the caller's capacity)
    int32_t* p;
    int32_t* end;
    bool overflow = false;
    IdSink& operator++(int) { return *this; }
    struct Ref {
        IdSink& s;
        void operator=(int32_t v) {
            if (s.p < s.end)
                *s.p++ = v;
            else
                s.overflow = true;
        }
    };
    Ref operator*() { return Ref{*this}; }
};
inline bool operator==(const IdSink& a, const IdSink& b) { return a.p == b.p; }

// Pre-tokenise + word-piece one NORMALISED segment (split on the table's spaces and punctuation).
inline void segment_to_ids(const asp_wordpiece& wp, const std::string& buf, size_t max_chars, IdSink& out, std::string& word) {
    size_t chars = 0;
    auto flush = [&]() {
        if (!word.empty()) {
            const IdSink first = out;
            wordpiece(wp, word.data(), word.size(), chars, max_chars, out);
            (void)first;
            word.clear();
            chars = 0;
        }
    };
    for (size_t i = 0; i < buf.size();) {
        uint32_t cp = 0;
        int len = utf8_decode(buf.data(), i, buf.size(), cp);
        if (!len) len = 1, cp = 0xFFFD;  // cannot happen: the blob holds the normaliser's own (valid) output
        const uint8_t cls = cp < 0x10000 ? wp.out_class[cp] : 0;
        if (cls == 1) {
            flush();
        } else if (cls == 2) {
            flush();
            wordpiece(wp, buf.data() + i, (size_t)len, 1, max_chars, out);
        } else {
            word.append(buf, i, (size_t)len);
            ++chars;
        }
        i += (size_t)len;
    }
    flush();
}

// Sentence with non-ASCII characters, through the per-code-point tables.  -1: leave it to the original tokenizer.
int64_t encode_sentence_unicode(const asp_wordpiece& wp, const char* s, size_t n, size_t max_chars, int32_t* out_base) {
    IdSink out{out_base, out_base + n};
    std::string buf, word;
    buf.reserve(n + 16);
    size_t i = 0;
    while (i < n) {
        const unsigned char c = (unsigned char)s[i];
        if (!wp.specials.empty() && c == (unsigned char)wp.specials.front().first[0]) {
            const std::pair<std::string, int32_t>* hit = nullptr;
            for (const auto& sp : wp.specials)
                if (sp.first.size() <= n - i && memcmp(s + i, sp.first.data(), sp.first.size()) == 0 &&
                    (!hit || sp.first.size() > hit->first.size()))
                    hit = &sp;
            if (hit) {
                segment_to_ids(wp, buf, max_chars, out, word);
                buf.clear();
                *out = hit->second;
                i += hit->first.size();
                continue;
            }
        }
        uint32_t cp = 0;
        const int len = utf8_decode(s, i, n, cp);
        if (!len || cp >= 0x10000 || wp.fallback[cp]) return -1;
        buf.append(wp.norm_blob, wp.norm_offsets[cp], wp.norm_offsets[cp + 1] - wp.norm_offsets[cp]);
        i += (size_t)len;
    }
    segment_to_ids(wp, buf, max_chars, out, word);
    return out.overflow ? -1 : out.p - out_base;
}

}  // namespace

extern "C" asp_wordpiece* asp_wordpiece_create(const char* vocab_blob, const int64_t* vocab_offsets, int n_vocab, int lower_case,
                                               int unk_id, const int32_t* special_ids, int n_special) {
    if (!vocab_blob || !vocab_offsets || n_vocab <= 0 || unk_id < 0 || unk_id >= n_vocab || (n_special && !special_ids)) {
        asp::set_error("asp_wordpiece_create: bad arguments (n_vocab %d, unk_id %d)", n_vocab, unk_id);
        return nullptr;
    }
    auto* wp = new asp_wordpiece;
    wp->blob.assign(vocab_blob, (size_t)vocab_offsets[n_vocab]);
    wp->unk_id = unk_id;
    wp->lower_case = lower_case != 0;
    std::vector<char> is_special((size_t)n_vocab, 0);
    for (int k = 0; k < n_special; ++k)
        if (special_ids[k] >= 0 && special_ids[k] < n_vocab) is_special[(size_t)special_ids[k]] = 1;
    wp->initial.reserve((size_t)n_vocab * 2);
    wp->continuation.reserve((size_t)n_vocab);
    for (int id = 0; id < n_vocab; ++id) {
        const std::string_view tok(wp->blob.data() + vocab_offsets[id], (size_t)(vocab_offsets[id + 1] - vocab_offsets[id]));
        if (tok.empty()) continue;
        if (is_special[(size_t)id]) wp->specials.emplace_back(std::string(tok), id);
        if (tok.size() > 2 && tok[0] == '#' && tok[1] == '#') {
            wp->continuation.emplace(tok.substr(2), id);
            wp->max_piece = std::max(wp->max_piece, tok.size() - 2);
        }
        // a "##x" entry is also reachable as a word-initial piece when the text itself reads "##x"... it never is after
        // punctuation splitting ('#' is punctuation), so only genuine word-initial entries go into `initial`
        else {
            wp->initial.emplace(tok, id);
            wp->max_piece = std::max(wp->max_piece, tok.size());
        }
    }
    // single-character punctuation pieces such as "#" are ordinary entries of `initial`; specials all start with the
    // same character in BERT vocabularies ('['), which encode_sentence uses as its cheap pre-test -- enforce it
    for (const auto& sp : wp->specials)
        if (sp.first[0] != wp->specials.front().first[0]) {
            asp::set_error("asp_wordpiece_create: special tokens must share their first character ('%s' vs '%s')",
                           sp.first.c_str(), wp->specials.front().first.c_str());
            delete wp;
            return nullptr;
        }
    return wp;
}

extern "C" void asp_wordpiece_destroy(asp_wordpiece* wp) { delete wp; }

extern "C" int asp_wordpiece_set_unicode(asp_wordpiece* wp, const uint32_t* norm_offsets, const char* norm_blob,
                                         const uint8_t* out_class, const uint8_t* fallback) {
    ASP_REQUIRE(wp && norm_offsets && norm_blob && out_class && fallback, "asp_wordpiece_set_unicode: NULL argument");
    // the tables are a function of the tokenizer: installed once, never replaced (other threads may be reading them)
    if (wp->unicode.load(std::memory_order_acquire)) return ASP_OK;
    for (int cp = 0; cp < 0x10000; ++cp)
        ASP_REQUIRE(norm_offsets[cp] <= norm_offsets[cp + 1], "asp_wordpiece_set_unicode: offsets must not decrease (at U+%04X)", cp);
    wp->norm_offsets.assign(norm_offsets, norm_offsets + 0x10001);
    wp->norm_blob.assign(norm_blob, norm_offsets[0x10000]);
    wp->out_class.assign(out_class, out_class + 0x10000);
    wp->fallback.assign(fallback, fallback + 0x10000);
    wp->unicode.store(true, std::memory_order_release);
    return ASP_OK;
}

extern "C" int asp_wordpiece_encode(const asp_wordpiece* wp, const char* text, const int64_t* offsets, int n_sent,
                                    int max_chars_per_word, int threads, int32_t* out_ids, int64_t* out_offsets,
                                    uint8_t* needs_fallback) {
    ASP_REQUIRE(wp && offsets && out_offsets && needs_fallback && n_sent >= 0, "asp_wordpiece_encode: NULL argument");
    ASP_REQUIRE(n_sent == 0 || (text && out_ids) || offsets[n_sent] == offsets[0], "asp_wordpiece_encode: NULL text/out_ids");
    ASP_REQUIRE(max_chars_per_word >= 1, "asp_wordpiece_encode: max_chars_per_word must be >= 1");
    out_offsets[0] = 0;
    if (n_sent == 0) return ASP_OK;
    for (int i = 0; i < n_sent; ++i)
        ASP_REQUIRE(offsets[i] <= offsets[i + 1], "asp_wordpiece_encode: offsets must not decrease (sentence %d)", i);
    // pass 1 (parallel): sentence i writes its ids at its own byte offset -- always enough room -- and its count
    std::vector<int64_t> counts((size_t)n_sent);
    const int64_t base = offsets[0];
    auto work = [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            const int64_t n = encode_sentence(*wp, text + offsets[i], (size_t)(offsets[i + 1] - offsets[i]),
                                              (size_t)max_chars_per_word, out_ids + (offsets[i] - base));
            needs_fallback[i] = n < 0;
            counts[(size_t)i] = n < 0 ? 0 : n;
        }
    };
    const int nt = std::max(1, std::min({threads, n_sent / 64 + 1, 64}));
    if (nt == 1) {
        work(0, n_sent);
    } else {
        std::vector<std::thread> pool;
        const int per = (n_sent + nt - 1) / nt;
        for (int t = 0; t < nt; ++t) pool.emplace_back(work, std::min(n_sent, t * per), std::min(n_sent, (t + 1) * per));
        for (auto& th : pool) th.join();
    }
    // pass 2: compact to the front (destination never overtakes the source)
    int64_t w = 0;
    for (int i = 0; i < n_sent; ++i) {
        const int64_t src = offsets[i] - base, n = counts[(size_t)i];
        if (n && src != w) memmove(out_ids + w, out_ids + src, (size_t)n * sizeof(int32_t));
        w += n;
        out_offsets[i + 1] = w;
    }
    return ASP_OK;
}

// ---- sequence assembly (examples/ex_aspire_consent.py:120-173) ----------------------------------------------------
// Document d owns doc_sents[d] consecutive "sentences" (element 0 is the title); sentence k's ids are
// ids[sent_offsets[k] .. sent_offsets[k+1]).  At most `budget` (500) word pieces are kept: the sentence that crosses
// the budget is cut to fit (kept if at least one piece fits) and later ones are dropped; the title is encoded but not
// pooled.  seq_lens counts [CLS] .. [SEP]; abs_lens counts the kept abstract sentences.
namespace {
template <typename F>
inline void walk_doc(const int64_t* sent_offsets, int64_t first, int n, int budget, F&& keep) {
    int used = 0;
    for (int k = 0; k < n; ++k) {
        const int len = (int)(sent_offsets[first + k + 1] - sent_offsets[first + k]);
        const int room = budget - used, take = std::min(len, room);
        if (take > 0 || len == 0) keep(k, used, take);
        if (len > room) break;
        used += take;
    }
}
}  // namespace

extern "C" int asp_abstracts_plan(const int64_t* sent_offsets, const int32_t* doc_sents, int n_docs, int budget,
                                  int32_t* seq_lens, int32_t* abs_lens) {
    ASP_REQUIRE(sent_offsets && doc_sents && seq_lens && abs_lens && n_docs >= 0 && budget >= 1, "asp_abstracts_plan: bad argument");
    int64_t first = 0;
    for (int d = 0; d < n_docs; ++d) {
        ASP_REQUIRE(doc_sents[d] >= 1, "asp_abstracts_plan: document %d has no title element", d);
        for (int k = 0; k < doc_sents[d]; ++k)
            ASP_REQUIRE(sent_offsets[first + k] <= sent_offsets[first + k + 1],
                        "asp_abstracts_plan: sentence offsets must not decrease (document %d, sentence %d)", d, k);
        int total = 0, spans = 0;
        walk_doc(sent_offsets, first, doc_sents[d], budget, [&](int k, int, int take) {
            total += take;
            spans += k > 0;
        });
        seq_lens[d] = total + 2;
        abs_lens[d] = spans;
        first += doc_sents[d];
    }
    return ASP_OK;
}

extern "C" int asp_abstracts_fill(const int32_t* ids, const int64_t* sent_offsets, const int32_t* doc_sents, int n_docs,
                                  int budget, int cls_id, int sep_id, int64_t pad_id, int width, int max_sents, int64_t* tokid,
                                  int64_t* seg, int64_t* attn, int32_t* spans) {
    ASP_REQUIRE(sent_offsets && doc_sents && tokid && seg && attn && spans && n_docs >= 0 && width >= 2 && max_sents >= 0,
                "asp_abstracts_fill: bad argument");
    int64_t first = 0;
    for (int d = 0; d < n_docs; ++d) {
        ASP_REQUIRE(doc_sents[d] >= 1, "asp_abstracts_fill: document %d has no title element", d);
        for (int k = 0; k < doc_sents[d]; ++k)
            ASP_REQUIRE(sent_offsets[first + k] <= sent_offsets[first + k + 1],
                        "asp_abstracts_fill: sentence offsets must not decrease (document %d, sentence %d)", d, k);
        int64_t* row = tokid + (size_t)d * width;
        int32_t* sp = spans + (size_t)d * max_sents * 2;
        for (int s = 0; s < 2 * max_sents; ++s) sp[s] = -1;
        int n = 1, rc = ASP_OK;
        row[0] = cls_id;
        walk_doc(sent_offsets, first, doc_sents[d], budget, [&](int k, int used, int take) {
            if (used + take + 2 > width || (k > 0 && k - 1 >= max_sents)) {
                rc = ASP_ERR_INVALID;
                return;
            }
            const int32_t* src = ids + sent_offsets[first + k];
            for (int j = 0; j < take; ++j) row[1 + used + j] = src[j];
            n = 1 + used + take;
            if (k > 0 && take > 0) {  // +1: the [CLS] in front; empty sentences keep the (-1, -1) "no span" marker
                sp[2 * (k - 1)] = used + 1;
                sp[2 * (k - 1) + 1] = used + 1 + take;
            }
        });
        ASP_REQUIRE(rc == ASP_OK, "asp_abstracts_fill: document %d does not fit width %d / max_sents %d (plan mismatch)", d, width,
                    max_sents);
        row[n++] = sep_id;
        for (int j = 0; j < width; ++j) {
            seg[(size_t)d * width + j] = j < n ? 0 : pad_id;
            attn[(size_t)d * width + j] = j < n ? 1 : pad_id;
            if (j >= n) row[j] = pad_id;
        }
        first += doc_sents[d];
    }
    return ASP_OK;
}
